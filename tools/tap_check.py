"""Parity of strided / non-square convolutions that take the per-tap TMA modes: python tools/tap_check.py"""
import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cobevt_b200 import ops
CASES = [  # N, H, W, Cin, Cout, (kh, kw), (sh, sw), (ph, pw)
    (2, 40, 48, 64, 128, (3, 3), (2, 2), (1, 1)),
    (2, 40, 48, 128, 256, (3, 3), (2, 2), (1, 1)),
    (2, 40, 48, 64, 128, (1, 1), (2, 2), (0, 0)),
    (2, 40, 48, 128, 256, (1, 1), (2, 2), (0, 0)),
    (2, 40, 48, 64, 128, (6, 3), (2, 1), (2, 1)),
    (2, 37, 50, 8, 64, (7, 4), (2, 1), (0, 0)),
    (2, 40, 48, 32, 32, (3, 3), (1, 1), (1, 1)),
    (1, 128, 64, 64, 128, (6, 3), (2, 1), (2, 1)),
]
for N, H, W, Cin, Cout, (kh, kw), st, pd in CASES:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, H, W, Cin, generator=g).half().cuda()
    w = torch.randn(Cout, Cin, kh, kw, generator=g) / math.sqrt(Cin * kh * kw)
    b = torch.randn(Cout, generator=g)
    pc = ops.PackedConv(w, b, "cuda")
    y, _, _ = ops.conv2d_nhwc(x, pc, stride=st, pad=pd, act=0)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.half().float().cuda(), b.cuda(), stride=st, padding=pd).permute(0, 2, 3, 1)
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
    print(f"Cin={Cin:4d} Cout={Cout:4d} k={kh}x{kw} s={st} p={pd} HxW={H}x{W}: relerr {err:.2e} {'OK' if err < 3e-3 else 'FAIL'}")
