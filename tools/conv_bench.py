"""Micro-benchmark of single conv_gemm launches (CUDA events): python tools/conv_bench.py"""
import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cobevt_b200 import ops

CASES = [  # name, N, H, W, Cin, Cout, k, stride, pad, res
    ("layer1 3x3 64->64 +res", 120, 56, 120, 64, 64, 3, 1, 1, True),
    ("layer1 3x3 64->64", 120, 56, 120, 64, 64, 3, 1, 1, False),
    ("layer2 3x3 128->128", 120, 28, 60, 128, 128, 3, 1, 1, False),
    ("layer3 3x3 256->256", 120, 14, 30, 256, 256, 3, 1, 1, False),
    ("linear K128 N128 M806400", 806400, 1, 1, 128, 128, 1, 1, 0, False),
    ("linear K128 N128 +res", 806400, 1, 1, 128, 128, 1, 1, 0, True),
    ("linear K64 N128", 806400, 1, 1, 64, 128, 1, 1, 0, False),
    ("linear K128 N128 M/4", 201600, 1, 1, 128, 128, 1, 1, 0, False),
    ("linear K128 N128 M/16", 50400, 1, 1, 128, 128, 1, 1, 0, False),
    ("linear K128 N128 M=128", 128, 1, 1, 128, 128, 1, 1, 0, False),
    ("stem 7x7 s2", 120, 224, 480, 8, 64, 7, 2, 3, False),
]
for name, N, H, W, Cin, Cout, k, st, pd, use_res in CASES:
    x = torch.randn(N, H, W, Cin, device="cuda").half()
    w = torch.randn(Cout, Cin, k, k) / math.sqrt(Cin * k * k)
    pc = ops.PackedConv(w, torch.zeros(Cout), "cuda", kw_pad=8 if k == 7 else None)
    y, _, _ = ops.conv2d_nhwc(x, pc, stride=st, pad=pd, act=1)
    res = torch.randn_like(y) if use_res else None
    for _ in range(3):
        ops.conv2d_nhwc(x, pc, stride=st, pad=pd, act=1, res=res, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        ops.conv2d_nhwc(x, pc, stride=st, pad=pd, act=1, res=res, out=y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2 * y.numel() // Cout * pc.k_alg * Cout
    print(f"{name:28s} {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TF/s  dbg={os.environ.get('COBEVT_CONV_DEBUG','0')}")
