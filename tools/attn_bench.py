"""Stage-0 FAX cross-attention alone (the dominant window_attn launch of the north_star step):
python tools/attn_bench.py [reps].  COBEVT_ATTN_DEBUG=<bits> selects the kernel's timing experiments."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cobevt_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B, heads, C = 20, 4, 128
g = torch.Generator().manual_seed(0)
q = torch.randn(B, 6, 128, 128, C, generator=g).half().cuda()
k = torch.randn(B, 6, 56, 120, C, generator=g).half().cuda()
v = torch.randn(B, 6, 56, 120, C, generator=g).half().cuda()
qm = ops.token_map(6, 128, 128, 16, 16, False)
km = ops.token_map(6, 56, 120, 7, 15, False)
for _ in range(2):
    ops.window_attention(q, k, v, qm, km, heads, 32 ** -0.5, True)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.window_attention(q, k, v, qm, km, heads, 32 ** -0.5, True); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ms = sorted(ts)[len(ts) // 2]
exps = B * 64 * heads * 1536 * 640
print(f"dbg={os.environ.get('COBEVT_ATTN_DEBUG', '0'):>2s}  {ms*1e3:8.1f} us   {exps/ms/1e9:6.3f} Texp/s")
