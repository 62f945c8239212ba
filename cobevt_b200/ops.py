"""Thin Python wrappers over the C ABI: torch tensors in, raw pointers across the boundary.

Tensors are device buffers only (torch is plumbing: allocation + streams).  Every wrapper
raises on a non-zero return code; nothing here computes on the host or falls back.
"""
import ctypes

import torch

from . import _lib
from ._lib import ConvParams, AttnParams, TokenMap

F16, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2


def dtype_code(dt):
    if dt == torch.float16:
        return F16
    if dt == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported 16-bit dtype {dt}")


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def pick_block_n(cout):
    """N tile used for a layer with `cout` output channels."""
    if cout <= 16:
        return 16
    if cout <= 32:
        return 32
    if cout <= 64:
        return 64
    if cout <= 128 or cout % 256 != 0:
        return 128
    return 256


class PackedConv:
    """Host-packed weights (+ fp32 bias) of one conv / linear layer, resident on the device.

    weight: (Cout, Cin, KH, KW) or (Cout, Cin) fp32 tensor (any device); bias: (Cout,) or None.
    BatchNorm folding and channel permutations are done by the caller before packing.
    """

    def __init__(self, weight, bias, device, dtype=torch.float16, cin_pad=None, block_n=None):
        w = weight.detach().float().cpu()
        if w.dim() == 2:
            w = w[:, :, None, None]
        cout, cin, kh, kw = w.shape
        cin_p = cin_pad if cin_pad is not None else (cin + 7) // 8 * 8
        if cin_p != cin:
            w = torch.cat([w, torch.zeros(cout, cin_p - cin, kh, kw)], dim=1)
        self.cout, self.cin, self.kh, self.kw = cout, cin_p, kh, kw
        self.K = kh * kw * cin_p
        self.block_n = block_n or pick_block_n(cout)
        self.dtype = dtype
        # [Cout][KH][KW][Cin] row-major, 16-bit
        w16 = w.permute(0, 2, 3, 1).contiguous().reshape(cout, self.K).to(dtype).contiguous()
        L = _lib.lib()
        nbytes = L.cobevt_conv_packed_bytes(cout, self.K, self.block_n)
        packed = torch.empty(nbytes // 2, dtype=torch.int16)
        src = w16.view(torch.int16)
        _lib.check(L.cobevt_conv_pack_weight(ctypes.c_void_p(src.data_ptr()), cout, self.K,
                                             self.block_n, ctypes.c_void_p(packed.data_ptr())),
                   "cobevt_conv_pack_weight")
        self.w = packed.to(device)
        self.bias = None if bias is None else bias.detach().float().contiguous().to(device)


def conv2d_nhwc(x, pc, stride=1, pad=0, act=ACT_NONE, res=None, res32=None, out=None,
                out32=False, want16=True, ln_out=False, in_upsample=False, ln_eps=1e-5):
    """x: (N, H, W, Cin) 16-bit contiguous.  Returns (y16 | None, y32 | None, y_ln | None)."""
    assert x.is_cuda and x.is_contiguous() and x.dim() == 4
    N, H, W, Cin = x.shape
    assert Cin == pc.cin, f"Cin {Cin} != packed {pc.cin}"
    Hl, Wl = (2 * H, 2 * W) if in_upsample else (H, W)
    OH = (Hl + 2 * pad - pc.kh) // stride + 1
    OW = (Wl + 2 * pad - pc.kw) // stride + 1
    M = N * OH * OW
    y = out if out is not None else (torch.empty((N, OH, OW, pc.cout), dtype=x.dtype, device=x.device) if want16 else None)
    y32 = torch.empty((N, OH, OW, pc.cout), dtype=torch.float32, device=x.device) if out32 else None
    yln = torch.empty((N, OH, OW, pc.cout), dtype=x.dtype, device=x.device) if ln_out else None
    if res is not None:
        assert res.is_contiguous() and res.numel() == M * pc.cout and res.dtype == x.dtype
    if res32 is not None:
        assert res32.is_contiguous() and res32.numel() == M * pc.cout and res32.dtype == torch.float32
    p = ConvParams()
    p.x, p.w, p.bias = _ptr(x), _ptr(pc.w), _ptr(pc.bias)
    p.res, p.res32, p.y, p.y32, p.y_ln = _ptr(res), _ptr(res32), _ptr(y), _ptr(y32), _ptr(yln)
    p.N, p.Hin, p.Win, p.Cin = N, H, W, Cin
    p.OH, p.OW, p.Cout = OH, OW, pc.cout
    p.KH, p.KW, p.stride, p.pad = pc.kh, pc.kw, stride, pad
    p.in_upsample = 1 if in_upsample else 0
    p.act, p.dtype, p.block_n, p.ln_eps = act, dtype_code(x.dtype), pc.block_n, ln_eps
    _lib.check(_lib.lib().cobevt_conv2d_nhwc_fwd(ctypes.byref(p), _stream()), "cobevt_conv2d_nhwc_fwd")
    return y, y32, yln


def linear_tokens(x, pc, **kw):
    """x: (..., Cin) 16-bit tokens -> (..., Cout); a 1x1 'conv' over a (rows,1,1,C) map."""
    shp = x.shape
    x4 = x.reshape(-1, 1, 1, shp[-1])
    for k in ("res", "res32"):
        if kw.get(k) is not None:
            kw[k] = kw[k].reshape(-1, 1, 1, pc.cout)
    y, y32, yln = conv2d_nhwc(x4, pc, **kw)
    f = lambda t: None if t is None else t.reshape(*shp[:-1], pc.cout)
    return f(y), f(y32), f(yln)


def token_map(views, H, W, w1, w2, grid, Hv=None, Wv=None):
    m = TokenMap()
    m.views, m.H, m.W = views, H, W
    m.Hv, m.Wv = (H if Hv is None else Hv), (W if Wv is None else Wv)
    m.w1, m.w2, m.grid = w1, w2, 1 if grid else 0
    return m


def window_attention(q, k, v, qmap, kvmap, heads, scale, q_mean, k_pad=None, v_pad=None,
                     bias=None, key_mask=None):
    """q: (B, qviews, Hq, Wq, C); k, v: (B, kviews, Hk, Wk, C) 16-bit contiguous (valid extents).

    Returns (B, Hq, Wq, C) if q_mean else (B, qviews, Hq, Wq, C).
    """
    assert q.is_cuda and q.is_contiguous() and k.is_contiguous() and v.is_contiguous()
    B, qv, Hq, Wq, C = q.shape
    assert C == heads * 32 and k.shape == v.shape and k.shape[-1] == C
    assert (qv, Hq, Wq) == (qmap.views, qmap.Hv, qmap.Wv)
    assert tuple(k.shape[1:4]) == (kvmap.views, kvmap.Hv, kvmap.Wv)
    out = torch.empty((B, Hq, Wq, C) if q_mean else (B, qv, Hq, Wq, C), dtype=q.dtype, device=q.device)
    p = AttnParams()
    p.q, p.k, p.v, p.out = _ptr(q), _ptr(k), _ptr(v), _ptr(out)
    p.k_pad, p.v_pad = _ptr(k_pad), _ptr(v_pad)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    if key_mask is not None:
        assert key_mask.dtype == torch.uint8 and key_mask.is_contiguous()
    p.bias, p.key_mask = _ptr(bias), _ptr(key_mask)
    p.qmap, p.kvmap = qmap, kvmap
    p.B, p.heads, p.q_mean, p.scale, p.dtype = B, heads, 1 if q_mean else 0, scale, dtype_code(q.dtype)
    _lib.check(_lib.lib().cobevt_window_attn_fwd(ctypes.byref(p), _stream()), "cobevt_window_attn_fwd")
    return out
