"""BASELINE config C5: FuseBEVT window/grid attention micro-benchmark.  One SwapFusionBlockMask
(window attention + FFN + grid attention + FFN) on x (1, N, 128, H, W), mask of ones, N = 1..8 agents:
w = 8 at H = W = 256 (the reference's window; w = 7 does not divide 256 -- the reference raises) and
w = 7 at H = W = 252.     python tools/fax_microbench.py"""
import copy, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cobevt_b200 as cb
from cobevt_b200 import synth, ops

def run(N, w, HW):
    args = {'input_dim': 128, 'mlp_dim': 256, 'agent_size': N, 'window_size': w, 'dim_head': 32, 'drop_out': 0.0,
            'depth': 1, 'mask': True}
    m = cb.SwapFusionEncoder(args).eval()
    m.load_state_dict(synth.synth_state_dict(m, seed=0))
    m = m.cuda()
    x = torch.randn(1, N, HW, HW, 128, device="cuda").half()
    mask = torch.ones(1, HW, HW, N, dtype=torch.uint8, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            m.run_nhwc(x, mask)
        torch.cuda.synchronize()
        ops.STATS.reset(); ops.STATS.profile = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.run_nhwc(x, mask)
        e1.record(); torch.cuda.synchronize()
        ops.STATS.profile = False
    fam = ops.STATS.family_ms()
    att_ms, att_fl = fam.get("window_attn", 0.0), ops.STATS.flops["window_attn"]
    T = N * w * w
    exps = 2 * 4 * (HW // w) ** 2 * T * T              # two attentions, 4 heads
    io_gb = 2 * 2 * N * HW * HW * 128 * 2 / 1e9         # q,k,v read + out write per attention (16-bit), x2 attentions ~ min traffic
    print(f"N={N} w={w} H=W={HW} T={T:4d} | block {e0.elapsed_time(e1):7.3f} ms | attention {att_ms:7.3f} ms "
          f"{att_fl/att_ms/1e9:7.1f} TF/s {exps/att_ms/1e9:6.2f} Texp/s (MUFU peak 4.65) {4*io_gb/att_ms*1e3/2:7.0f} GB/s qkv+out | "
          f"gemm {fam.get('conv_gemm',0):6.3f} ms eltwise {fam.get('eltwise',0):6.3f} ms")

for N in range(1, 9):
    run(N, 8, 256)
for N in (1, 5, 8):
    run(N, 7, 252)
