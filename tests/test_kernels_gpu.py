"""GPU parity of the two tensor-core kernels against plain PyTorch fp32 math.

Inputs are rounded to the 16-bit storage type first, the reference then runs in fp32, so
the comparison isolates kernel arithmetic (fp32 accumulation, one final rounding).
Tolerance: max|err| <= 2e-3 * max|ref| for fp16 (1e-3 target + one output rounding),
1.5e-2 for bf16 (8-bit mantissa output rounding alone is 3.9e-3).
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {torch.float16: 2e-3, torch.bfloat16: 1.5e-2}


def relerr(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


@pytest.mark.parametrize("cfg", [
    # N, H, W, Cin, Cout, k, stride, pad, act, res, dtype
    (2, 20, 24, 64, 64, 3, 1, 1, 1, True, torch.float16),
    (2, 20, 24, 64, 128, 3, 2, 1, 1, False, torch.float16),
    (1, 33, 17, 128, 256, 1, 1, 0, 0, False, torch.float16),
    (1, 16, 16, 256, 512, 3, 1, 1, 1, True, torch.float16),
    (2, 30, 30, 8, 64, 7, 2, 3, 1, False, torch.float16),
    (1, 14, 14, 128, 32, 1, 1, 0, 1, False, torch.float16),
    (1, 14, 14, 32, 32, 3, 1, 1, 1, False, torch.float16),
    (3, 9, 9, 32, 2, 3, 1, 1, 0, False, torch.float16),
    (1, 40, 40, 128, 384, 1, 1, 0, 0, False, torch.float16),
    (2, 20, 24, 64, 64, 3, 1, 1, 2, True, torch.bfloat16),
    (1, 12, 12, 512, 512, 3, 1, 1, 1, False, torch.float16),
])
def test_conv_parity(cfg):
    from cobevt_b200 import ops
    N, H, W, Cin, Cout, k, stride, pad, act, use_res, dt = cfg
    g = torch.Generator().manual_seed(hash(cfg[:8]) % 1000)
    x = torch.randn(N, H, W, Cin, generator=g).to(dt)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)).to(dt)
    b = torch.randn(Cout, generator=g)
    pc = ops.PackedConv(w.float(), b, "cuda", dtype=dt)
    xd = x.cuda()
    ref = F.conv2d(xd.float().permute(0, 3, 1, 2), w.cuda().float(), b.cuda(), stride=stride, padding=pad)
    ref = ref.permute(0, 2, 3, 1).contiguous()
    res = None
    if use_res:
        res = torch.randn(ref.shape, generator=g).to(dt).cuda()
        ref = ref + res.float()
    if act == 1:
        ref = F.relu(ref)
    elif act == 2:
        ref = F.gelu(ref)
    want32 = Cout == 2
    y, y32, _ = ops.conv2d_nhwc(xd, pc, stride=stride, pad=pad, act=act, res=res, out32=want32)
    torch.cuda.synchronize()
    assert y.shape == ref.shape
    assert relerr(y, ref) < TOL[dt], relerr(y, ref)
    if want32:
        assert relerr(y32, ref) < 1e-4


def test_conv_ln_and_fp32_residual():
    from cobevt_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(700, 256, generator=g).half()
    w = (torch.randn(128, 256, generator=g) / 16).half()
    b = torch.randn(128, generator=g)
    r32 = torch.randn(700, 128, generator=g)
    pc = ops.PackedConv(w.float(), b, "cuda")
    y, y32, yln = ops.linear_tokens(x.cuda(), pc, res32=r32.cuda(), out32=True, ln_out=True)
    ref = x.cuda().float() @ w.cuda().float().t() + b.cuda() + r32.cuda()
    assert relerr(y32, ref) < 1e-4
    assert relerr(y, ref) < 2e-3
    refln = F.layer_norm(ref, (128,))
    assert relerr(yln, refln) < 2e-3


def test_conv_upsampled_input():
    from cobevt_b200 import ops
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 8, 8, 64, generator=g).half()
    w = (torch.randn(64, 64, 3, 3, generator=g) / 24).half()
    pc = ops.PackedConv(w.float(), None, "cuda")
    y, _, _ = ops.conv2d_nhwc(x.cuda(), pc, stride=1, pad=1, in_upsample=True)
    xu = F.interpolate(x.cuda().float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    ref = F.conv2d(xu, w.cuda().float(), padding=1).permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    assert relerr(y, ref) < 2e-3



@pytest.mark.parametrize("cfg", [
    # N, H, W, Cin, Cout, (kh, kw), (sh, sw), (ph, pw)   -- the per-tap / patch TMA producer modes
    (2, 40, 48, 64, 128, (3, 3), (2, 2), (1, 1)),
    (2, 40, 48, 128, 256, (3, 3), (2, 2), (1, 1)),      # two 64-channel slabs per tap
    (2, 40, 48, 128, 256, (1, 1), (2, 2), (0, 0)),
    (2, 40, 48, 64, 128, (6, 3), (2, 1), (2, 1)),       # PixelUnshuffle-folded downsample conv
    (2, 37, 50, 8, 64, (7, 4), (2, 1), (0, 0)),         # stem on RGBx super-pixels: in-place im2col (PATCH)
    (2, 30, 30, 8, 64, (7, 8), (2, 2), (3, 3)),         # 8 channels, strided along W: per-tap core-matrix columns
    (2, 40, 48, 32, 32, (3, 3), (1, 1), (1, 1)),        # 32 channels: SWIZZLE_64B tap tiles
])
def test_conv_strided_tap_modes(cfg):
    from cobevt_b200 import ops
    N, H, W, Cin, Cout, (kh, kw), st, pd = cfg
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, H, W, Cin, generator=g).half().cuda()
    w = torch.randn(Cout, Cin, kh, kw, generator=g) / math.sqrt(Cin * kh * kw)
    b = torch.randn(Cout, generator=g)
    pc = ops.PackedConv(w, b, "cuda")
    y, _, _ = ops.conv2d_nhwc(x, pc, stride=st, pad=pd, act=0)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.half().float().cuda(), b.cuda(),
                                     stride=st, padding=pd).permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    assert relerr(y, ref) < 2e-3, relerr(y, ref)

# ----------------------------------------------------------------------------- attention
def partition(t, w1, w2, grid):
    """(B, n, H, W, C) -> (B, windows, n*w1*w2, C) with the reference's einops patterns."""
    B, n, H, W, C = t.shape
    if grid:   # 'b n (w1 x) (w2 y) d -> b (x y) (n w1 w2) d'
        X, Y = H // w1, W // w2
        t = t.reshape(B, n, w1, X, w2, Y, C).permute(0, 3, 5, 1, 2, 4, 6)
    else:      # 'b n (x w1) (y w2) d -> b (x y) (n w1 w2) d'
        X, Y = H // w1, W // w2
        t = t.reshape(B, n, X, w1, Y, w2, C).permute(0, 2, 4, 1, 3, 5, 6)
    return t.reshape(B, X * Y, n * w1 * w2, C)


def ref_attention(q, k, v, qwin, kwin, qgrid, kgrid, heads, scale, q_mean, bias=None, mask=None):
    B, nq, Hq, Wq, C = q.shape
    qp = partition(q.float(), *qwin, qgrid)
    kp = partition(k.float(), *kwin, kgrid)
    vp = partition(v.float(), *kwin, kgrid)
    Wn, Lq, _ = qp.shape[1:]
    Lk = kp.shape[2]
    qh = qp.reshape(B, Wn, Lq, heads, 32).permute(0, 1, 3, 2, 4)
    kh = kp.reshape(B, Wn, Lk, heads, 32).permute(0, 1, 3, 2, 4)
    vh = vp.reshape(B, Wn, Lk, heads, 32).permute(0, 1, 3, 2, 4)
    s = scale * qh @ kh.transpose(-1, -2)
    if bias is not None:
        s = s + bias[None, None]
    if mask is not None:   # mask: (B, H, W, n) -> per window keys
        mp = partition(mask.permute(0, 3, 1, 2)[..., None].float(), *kwin, kgrid)[..., 0]  # B, Wn, Lk
        s = s.masked_fill(mp[:, :, None, None, :] == 0, float("-inf"))
    o = s.softmax(-1) @ vh                                   # B Wn h Lq 32
    o = o.permute(0, 1, 3, 2, 4).reshape(B, Wn, Lq, C)
    # un-partition queries
    w1, w2 = qwin
    X, Y = Hq // w1, Wq // w2
    o = o.reshape(B, X, Y, nq, w1, w2, C)
    if qgrid:
        o = o.permute(0, 3, 4, 1, 5, 2, 6).reshape(B, nq, Hq, Wq, C)
    else:
        o = o.permute(0, 3, 1, 4, 2, 5, 6).reshape(B, nq, Hq, Wq, C)
    return o.mean(1) if q_mean else o


@pytest.mark.parametrize("cfg", [
    # B, nq, nk, Hq, Wq, qwin, Hk, Wk, kwin, kgrid, q_mean, dtype
    (2, 4, 4, 32, 32, (16, 16), 16, 16, (8, 8), False, True, torch.float16),
    (2, 1, 4, 32, 32, (16, 16), 16, 16, (8, 8), True, True, torch.float16),
    (1, 6, 6, 32, 32, (16, 16), 14, 30, (7, 15), False, True, torch.float16),
    (1, 1, 6, 32, 32, (16, 16), 14, 30, (7, 15), True, True, torch.float16),
    (1, 1, 3, 32, 32, (32, 32), 16, 16, (16, 16), True, True, torch.bfloat16),
    # exactly 1024 resident keys: single Q buffer variant (no room for the prefetch buffer)
    (1, 2, 4, 32, 32, (16, 16), 32, 32, (16, 16), False, True, torch.float16),
    # > 1024 keys per window: the streaming (non-resident) kernel
    (1, 2, 3, 24, 24, (24, 24), 20, 20, (20, 20), False, True, torch.float16),
    (1, 1, 5, 16, 16, (16, 16), 15, 15, (15, 15), True, False, torch.float16),
])
def test_cross_attention_parity(cfg):
    from cobevt_b200 import ops
    B, nq, nk, Hq, Wq, qwin, Hk, Wk, kwin, kgrid, q_mean, dt = cfg
    heads, C = 4, 128
    g = torch.Generator().manual_seed(11)
    q = torch.randn(B, nq, Hq, Wq, C, generator=g).to(dt).cuda()
    k = torch.randn(B, nk, Hk, Wk, C, generator=g).to(dt).cuda()
    v = torch.randn(B, nk, Hk, Wk, C, generator=g).to(dt).cuda()
    scale = 32 ** -0.5
    qm = ops.token_map(nq, Hq, Wq, *qwin, False)
    km = ops.token_map(nk, Hk, Wk, *kwin, kgrid)
    out = ops.window_attention(q, k, v, qm, km, heads, scale, q_mean)
    ref = ref_attention(q, k, v, qwin, kwin, False, kgrid, heads, scale, q_mean)
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    assert relerr(out, ref) < TOL[dt], relerr(out, ref)


def test_cross_attention_padded_keys():
    """Padded key rows are constant tokens that take part in the softmax (quirk Q2)."""
    from cobevt_b200 import ops
    B, n, heads, C = 1, 2, 2, 64
    Hv, Wv, H, W = 10, 20, 12, 24          # stored 10x20, padded to 12x24, windows 6x12
    g = torch.Generator().manual_seed(3)
    q = torch.randn(B, 1, 20, 20, C, generator=g).half().cuda()
    k = torch.randn(B, n, Hv, Wv, C, generator=g).half().cuda()
    v = torch.randn(B, n, Hv, Wv, C, generator=g).half().cuda()
    kp = torch.randn(C, generator=g).half().cuda()
    vp = torch.randn(C, generator=g).half().cuda()
    scale = 32 ** -0.5
    for grid in (False, True):
        qm = ops.token_map(1, 20, 20, 10, 10, False)
        km = ops.token_map(n, H, W, 6, 12, grid, Hv=Hv, Wv=Wv)
        out = ops.window_attention(q, k, v, qm, km, heads, scale, True, k_pad=kp, v_pad=vp)
        kfull = kp.expand(B, n, H, W, C).clone(); kfull[:, :, :Hv, :Wv] = k
        vfull = vp.expand(B, n, H, W, C).clone(); vfull[:, :, :Hv, :Wv] = v
        ref = ref_attention(q, kfull, vfull, (10, 10), (6, 12), False, grid, heads, scale, True)
        assert relerr(out, ref) < 2e-3, (grid, relerr(out, ref))


@pytest.mark.parametrize("grid", [False, True])
@pytest.mark.parametrize("L", [1, 3, 5])
def test_self_attention_bias_mask(grid, L):
    """FuseBEVT-style: agents are extra token rows, rel-pos bias + key mask."""
    from cobevt_b200 import ops
    B, H, W, heads, C, w = 2, 16, 16, 4, 128, 8
    g = torch.Generator().manual_seed(7 + L)
    x = torch.randn(B, L, H, W, 3 * C, generator=g).half().cuda()
    q, k, v = [t.contiguous() for t in x.chunk(3, -1)]
    T = L * w * w
    bias = torch.randn(heads, T, T, generator=g).cuda()
    mask = (torch.rand(B, H, W, L, generator=g) > 0.3).to(torch.uint8)
    mask[..., 0] = 1
    mask = mask.cuda()
    scale = 32 ** -0.5
    m = ops.token_map(L, H, W, w, w, grid)
    out = ops.window_attention(q, k, v, m, m, heads, scale, False, bias=bias, key_mask=mask)
    ref = ref_attention(q, k, v, (w, w), (w, w), grid, grid, heads, scale, False, bias=bias, mask=mask)
    assert relerr(out, ref) < 2e-3, relerr(out, ref)
