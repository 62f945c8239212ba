"""Thin Python wrappers over the C ABI: torch tensors in, raw pointers across the boundary.

Tensors are device buffers only (torch is plumbing: allocation + streams).  Every wrapper
raises on a non-zero return code; nothing here computes on the host or falls back.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import ConvParams, AttnParams, TokenMap, MlpParams

F16, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU, ACT_SWISH = 0, 1, 2, 3


class _Stats:
    """Launch / algorithmic-FLOP accounting per kernel family, and (profile mode) CUDA-event
    timing of every launch on the launching stream.  bench.py reads this; nothing else does."""

    def __init__(self):
        self.reset()
        self.profile = False

    def reset(self):
        self.launches = {"conv_gemm": 0, "window_attn": 0, "eltwise": 0}
        self.flops = {"conv_gemm": 0, "window_attn": 0, "eltwise": 0}
        self.events = []
        self.detail = []

    def total_launches(self):
        return sum(self.launches.values())

    def family_ms(self):
        out = {}
        for fam, a, b in self.events:
            out[fam] = out.get(fam, 0.0) + a.elapsed_time(b)
        return out


STATS = _Stats()


class _timed:
    def __init__(self, fam, flops=0, label=""):
        self.fam, self.flops, self.label = fam, flops, label

    def __enter__(self):
        STATS.launches[self.fam] += 1
        STATS.flops[self.fam] += self.flops
        if STATS.profile:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if STATS.profile:
            self.b.record()
            STATS.events.append((self.fam, self.a, self.b))
            STATS.detail.append((self.fam, self.label, self.flops, self.a, self.b))
        return False


def dtype_code(dt):
    if dt == torch.float16:
        return F16
    if dt == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported 16-bit dtype {dt}")


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    """Current stream of the CURRENT device; every wrapper is entered with tensors of that device (checked by
    conv2d_nhwc / window_attention, the two kernels with per-device state)."""
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _same_device(t):
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"tensor on {t.device} but the current CUDA device is {torch.cuda.current_device()}: "
                           "wrap the call in torch.cuda.device(tensor.device)")


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

    def __init__(self, weight, bias, device, dtype=torch.float16, cin_pad=None, block_n=None, kw_pad=None):
        w = weight.detach().float().cpu()
        if w.dim() == 2:
            w = w[:, :, None, None]
        cout, cin, kh, kw = w.shape
        cin_p = cin_pad if cin_pad is not None else (cin + 7) // 8 * 8
        if cin_p != cin:
            w = torch.cat([w, torch.zeros(cout, cin_p - cin, kh, kw)], dim=1)
        self.k_alg = kh * kw * cin            # un-padded reduction length (algorithmic FLOPs)
        self.kw_logical = kw                  # output extent is computed from the true kernel width
        if kw_pad is not None and kw_pad > kw:   # zero taps appended on the right (keeps k-blocks row-aligned)
            w = torch.cat([w, torch.zeros(cout, w.shape[1], kh, kw_pad - kw)], dim=3)
            kw = kw_pad
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


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def conv2d_nhwc(x, pc, stride=1, pad=0, act=ACT_NONE, res=None, res32=None, out=None,
                out32=False, want16=True, ln_out=False, ln_gamma=None, ln_beta=None,
                in_upsample=False, ln_eps=1e-5, res_rows=0, out_hw=None):
    """x: (N, H, W, Cin) 16-bit contiguous.  Returns (y16 | None, y32 | None, y_ln | None)."""
    assert x.is_cuda and x.is_contiguous() and x.dim() == 4
    _same_device(x)
    N, H, W, Cin = x.shape
    assert Cin == pc.cin, f"Cin {Cin} != packed {pc.cin}"
    Hl, Wl = (2 * H, 2 * W) if in_upsample else (H, W)
    (sh, sw), (ph, pw) = _pair(stride), _pair(pad)
    OH = (Hl + 2 * ph - pc.kh) // sh + 1
    OW = (Wl + 2 * pw - pc.kw_logical) // sw + 1
    if out_hw is not None:           # caller-chosen output window (must lie inside the computed one)
        assert out_hw[0] <= OH and out_hw[1] <= OW
        OH, OW = out_hw
    M = N * OH * OW
    y = out if out is not None else (torch.empty((N, OH, OW, pc.cout), dtype=x.dtype, device=x.device) if want16 else None)
    y32 = torch.empty((N, OH, OW, pc.cout), dtype=torch.float32, device=x.device) if out32 else None
    yln = torch.empty((N, OH, OW, pc.cout), dtype=x.dtype, device=x.device) if ln_out else None
    rrows = res_rows if res_rows else M
    if res is not None:
        assert res.is_contiguous() and res.numel() == rrows * pc.cout and res.dtype == x.dtype
    if res32 is not None:
        assert res32.is_contiguous() and res32.numel() == rrows * pc.cout and res32.dtype == torch.float32
    p = ConvParams()
    p.x, p.w, p.bias = _ptr(x), _ptr(pc.w), _ptr(pc.bias)
    p.res, p.res32, p.y, p.y32, p.y_ln = _ptr(res), _ptr(res32), _ptr(y), _ptr(y32), _ptr(yln)
    p.N, p.Hin, p.Win, p.Cin = N, H, W, Cin
    p.OH, p.OW, p.Cout = OH, OW, pc.cout
    p.KH, p.KW = pc.kh, pc.kw
    p.stride_h, p.stride_w, p.pad_h, p.pad_w = sh, sw, ph, pw
    p.res_rows = res_rows
    p.ln_gamma, p.ln_beta = _ptr(ln_gamma), _ptr(ln_beta)
    p.in_upsample = 1 if in_upsample else 0
    p.act, p.dtype, p.block_n, p.ln_eps = act, dtype_code(x.dtype), pc.block_n, ln_eps
    with _timed("conv_gemm", 2 * M * pc.k_alg * pc.cout,
                f"M={M} K={pc.K} N={pc.cout} k{pc.kh}x{pc.kw} s{sh} bn{pc.block_n} in{H}x{W}x{Cin}"):
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


class PackedMlp:
    """Weights of one PreNorm feed-forward (LayerNorm affine already folded into w1 / b1), packed for the fused
    token-MLP kernel when it is the 128 -> 256 -> 128 shape that kernel is built for, and for the two-GEMM path
    otherwise."""

    def __init__(self, w1, b1, w2, b2, device, dtype):
        self.fused = tuple(w1.shape) == (256, 128) and tuple(w2.shape) == (128, 256)
        self.dtype = dtype
        if self.fused:
            self.a = PackedConv(w1, b1, device, dtype, block_n=128)
            self.b = PackedConv(w2, b2, device, dtype, block_n=128)
        else:
            self.a = PackedConv(w1, b1, device, dtype)
            self.b = PackedConv(w2, b2, device, dtype)


def token_mlp(x_ln, mlp, res, want16=True, ln_out=True, ln_gamma=None, ln_beta=None, ln_eps=1e-5):
    """x_ln (..., C) plain-normalised tokens, res (..., C) residual stream -> (y | None, y_ln | None):
    y = res + W2 gelu(W1 x_ln + b1) + b2, y_ln = LayerNorm(y)."""
    if not mlp.fused:
        hmid = linear_tokens(x_ln, mlp.a, act=ACT_GELU)[0]
        y, _, yln = linear_tokens(hmid, mlp.b, res=res, want16=want16, ln_out=ln_out, ln_gamma=ln_gamma, ln_beta=ln_beta)
        return y, yln
    shp = x_ln.shape
    C = shp[-1]
    M = x_ln.numel() // C
    assert x_ln.is_cuda and x_ln.is_contiguous() and res.is_contiguous() and res.numel() == M * C and res.dtype == x_ln.dtype
    _same_device(x_ln)
    y = torch.empty(shp, dtype=x_ln.dtype, device=x_ln.device) if want16 else None
    yln = torch.empty(shp, dtype=x_ln.dtype, device=x_ln.device) if ln_out else None
    p = MlpParams()
    p.x_ln, p.w1, p.b1, p.w2, p.b2 = _ptr(x_ln), _ptr(mlp.a.w), _ptr(mlp.a.bias), _ptr(mlp.b.w), _ptr(mlp.b.bias)
    p.res, p.y, p.y_ln, p.ln_gamma, p.ln_beta = _ptr(res), _ptr(y), _ptr(yln), _ptr(ln_gamma), _ptr(ln_beta)
    p.M, p.C, p.H, p.dtype, p.ln_eps = M, C, 256, dtype_code(x_ln.dtype), ln_eps
    with _timed("conv_gemm", 2 * M * (128 * 256 + 256 * 128), f"token_mlp M={M} 128->256->128"):
        _lib.check(_lib.lib().cobevt_token_mlp_fwd(ctypes.byref(p), _stream()), "cobevt_token_mlp_fwd")
    return y, yln


def token_map(views, H, W, w1, w2, grid, Hv=None, Wv=None):
    m = TokenMap()
    m.views, m.H, m.W = views, H, W
    m.Hv, m.Wv = (H if Hv is None else Hv), (W if Wv is None else Wv)
    m.w1, m.w2, m.grid = w1, w2, 1 if grid else 0
    return m




_RESIDENT_KEYS = None


def _resident_keys():
    """Keys one CTA of the four-warpgroup attention kernel keeps resident (asked once from the library)."""
    global _RESIDENT_KEYS
    if _RESIDENT_KEYS is None:
        _RESIDENT_KEYS = int(_lib.lib().cobevt_attn_resident_keys())
    return _RESIDENT_KEYS


class AttnBias:
    """A constant additive attention bias (heads, Lq, Lk) fp32, held in both layouts the kernels read: row-major, and
    key-group-major (heads, Lk/4, Lq, 4) for the resident-key kernel (a warp's 32 query rows then read 512 contiguous
    bytes per key group).  Built once per plan; `window_attention` picks the layout."""

    def __init__(self, bias):
        assert bias.dtype == torch.float32 and bias.dim() == 3
        self.raw = bias.contiguous()
        h, lq, lk = bias.shape
        self.kq = None
        if lk % 4 == 0 and os.environ.get("COBEVT_ATTN_BIAS_KQ", "1") != "0":      # =0: row-major only (A/B timing)
            self.kq = self.raw.view(h, lq, lk // 4, 4).permute(0, 2, 1, 3).contiguous()


def window_attention(q, k, v, qmap, kvmap, heads, scale, q_mean, k_pad=None, v_pad=None,
                     bias=None, key_mask=None):
    """q: (B, qviews, Hq, Wq, C); k, v: (B, kviews, Hk, Wk, C) 16-bit contiguous (valid extents).

    Returns (B, Hq, Wq, C) if q_mean else (B, qviews, Hq, Wq, C).
    """
    C = heads * 32
    B, qv, Hq, Wq, _ = q.shape
    _same_device(q)

    def tok_stride(t):
        # t may be a last-dim slice of a wider contiguous token tensor
        assert t.is_cuda and t.shape[-1] == C and t.stride(-1) == 1
        ts = t.stride(-2)
        exp = ts
        for d in range(t.dim() - 2, -1, -1):
            assert t.stride(d) == exp or t.shape[d] == 1, "tokens must be densely strided"
            exp *= t.shape[d]
        return ts

    qs, ks, vs = tok_stride(q), tok_stride(k), tok_stride(v)
    assert ks == vs and k.shape == v.shape
    assert (qv, Hq, Wq) == (qmap.views, qmap.Hv, qmap.Wv)
    assert tuple(k.shape[1:4]) == (kvmap.views, kvmap.Hv, kvmap.Wv)
    out = torch.empty((B, Hq, Wq, C) if q_mean else (B, qv, Hq, Wq, C), dtype=q.dtype, device=q.device)
    n_win = (qmap.H // qmap.w1) * (qmap.W // qmap.w2)
    lq, lk = qmap.views * qmap.w1 * qmap.w2, kvmap.views * kvmap.w1 * kvmap.w2
    p = AttnParams()
    p.q, p.k, p.v, p.out = _ptr(q), _ptr(k), _ptr(v), _ptr(out)
    p.k_pad, p.v_pad = _ptr(k_pad), _ptr(v_pad)
    bias_kq = None
    if isinstance(bias, AttnBias):
        bias, bias_kq = bias.raw, bias.kq
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    if key_mask is not None:
        assert key_mask.dtype == torch.uint8 and key_mask.is_contiguous()
    p.bias, p.key_mask = _ptr(bias), _ptr(key_mask)
    p.qmap, p.kvmap = qmap, kvmap
    p.B, p.heads, p.q_mean, p.scale, p.dtype = B, heads, 1 if q_mean else 0, scale, dtype_code(q.dtype)
    p.q_stride, p.kv_stride, p.out_stride = qs, ks, C
    label = f"B={B} win={n_win} heads={heads} Lq={lq} Lk={lk} mean={int(bool(q_mean))} bias={bias is not None}"

    def use_kq(pp, keys_per_call, keys_per_view):
        # the key-group-major copy serves the resident-key kernel only (see include/cobevt_b200.h, bias_layout)
        if bias_kq is not None and keys_per_call <= _resident_keys() and keys_per_view % 4 == 0:
            pp.bias, pp.bias_layout = _ptr(bias_kq), 1

    # windows whose keys do not fit one resident pass: one call per group of K/V views (each a resident segment
    # of the four-warpgroup kernel) + a log-sum-exp merge; e.g. the stage-2 global cross attention, 6 x 420 keys
    per_view = kvmap.w1 * kvmap.w2
    RESIDENT_KEYS = _resident_keys()
    if (lk > RESIDENT_KEYS and kvmap.views == 1 and not kvmap.grid and (kvmap.w1, kvmap.w2) == (kvmap.H, kvmap.W)
            and kvmap.H % 2 == 0 and per_view // 2 <= RESIDENT_KEYS and kvmap.Hv == kvmap.H and kvmap.Wv == kvmap.W):
        # one plane that IS the window (the 1024-token global self-attention, fax_modules.py:132-176): its rows split
        # into two half-height "views" -- the same memory and the same key order -- so each half is a resident segment
        # (only the token map changes: pointers and the token stride already describe both halves)
        kvmap = token_map(2, kvmap.H // 2, kvmap.W, kvmap.H // 2, kvmap.W, False)
        per_view = kvmap.w1 * kvmap.w2
    if (lk > RESIDENT_KEYS and kvmap.views > 1 and per_view <= RESIDENT_KEYS and lq > 2 * 128
            and key_mask is None and k_pad is None and kvmap.Hv == kvmap.H and kvmap.Wv == kvmap.W):
        vps = max(1, RESIDENT_KEYS // per_view)                  # views per segment
        while kvmap.views % vps:                                 # equal segments -> ONE launch covers them all
            vps -= 1
        nseg = kvmap.views // vps
        ps = AttnParams.from_buffer_copy(p)
        ps.kvmap = token_map(vps, kvmap.H, kvmap.W, kvmap.w1, kvmap.w2, bool(kvmap.grid))
        ps.out = None
        ps.kv_view0, ps.kv_views_total, ps.kv_segments = 0, kvmap.views, nseg
        use_kq(ps, vps * per_view, per_view)
        nbytes = int(_lib.lib().cobevt_attn_workspace_bytes(ctypes.byref(ps)))      # the library sizes its own scratch
        assert nbytes == nseg * B * n_win * heads * lq * 36 * 4, nbytes
        part = torch.empty((nseg, nbytes // (4 * nseg)), dtype=torch.float32, device=q.device)
        with _timed("window_attn", 4 * B * n_win * heads * lq * lk * 32, label + f" segs={nseg}"):
            ps.part = ctypes.c_void_p(part.data_ptr())
            ps.part_segment_stride = part[0].numel()
            _lib.check(_lib.lib().cobevt_window_attn_fwd(ctypes.byref(ps), _stream()), "cobevt_window_attn_fwd")
            _lib.check(_lib.lib().cobevt_attn_merge(ctypes.c_void_p(part.data_ptr()), nseg, part[0].numel(), _ptr(out),
                                                    ctypes.byref(qmap), B, heads, 1 if q_mean else 0, C,
                                                    dtype_code(q.dtype), _stream()), "cobevt_attn_merge")
        return out
    use_kq(p, lk, per_view)
    with _timed("window_attn", 4 * B * n_win * heads * lq * lk * 32, label):
        _lib.check(_lib.lib().cobevt_window_attn_fwd(ctypes.byref(p), _stream()), "cobevt_window_attn_fwd")
    return out


# --------------------------------------------------------------------------- element-wise ops
def _f32(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
    return ctypes.c_void_p(t.data_ptr())


def pack_rgb8(x, dtype):
    """x (..., 3) fp32 -> (..., 8) 16-bit (channels 3..7 zero)."""
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[-1] == 3
    y = torch.empty(*x.shape[:-1], 8, dtype=dtype, device=x.device)
    npix = x.numel() // 3
    with _timed("eltwise", 0, "pack_rgb8"):
        _lib.check(_lib.lib().cobevt_pack_rgb8(_ptr(x), _ptr(y), npix, dtype_code(dtype), _stream()), "cobevt_pack_rgb8")
    return y


def pack_rgb4_padded(x, dtype, pad_top, pad_left, Hp, Wp, nchw=False, mean=None, std=None, bgr=False):
    """x (N, H, W, 3) [or (N, 3, H, W) with nchw=True] fp32 -> (N, Hp, Wp, 4) 16-bit, zero border;
    optional per-channel (x - mean) / std (host triples).  uint8 (N, H, W, 3) frames: ((x / 255) - mean) / std
    with an optional BGR -> RGB swap (rgb_preprocessor.py:16-46)."""
    assert x.is_contiguous()
    if x.dtype == torch.uint8:
        assert not nchw and x.shape[-1] == 3
        N, H, W, _ = x.shape
        y = torch.empty(N, Hp, Wp, 4, dtype=dtype, device=x.device)
        m3 = s3 = None
        if mean is not None:
            m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
            s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
        with _timed("eltwise", 0, "stage_rgb_u8"):
            _lib.check(_lib.lib().cobevt_stage_rgb_u8(_ptr(x), _ptr(y), N, H, W, Hp, Wp, pad_top, pad_left, 1 if bgr else 0,
                                                      m3, s3, dtype_code(dtype), _stream()), "cobevt_stage_rgb_u8")
        return y
    assert x.dtype == torch.float32 and not bgr
    if nchw:
        N, _, H, W = x.shape
        assert x.shape[1] == 3
    else:
        N, H, W, _ = x.shape
        assert x.shape[-1] == 3
    y = torch.empty(N, Hp, Wp, 4, dtype=dtype, device=x.device)
    m3 = s3 = None
    if mean is not None:
        m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
        s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    with _timed("eltwise", 0, "pack_rgb4_padded"):
        _lib.check(_lib.lib().cobevt_pack_rgb4_padded(_ptr(x), _ptr(y), N, H, W, Hp, Wp, pad_top, pad_left,
                                                      1 if nchw else 0, m3, s3, dtype_code(dtype), _stream()),
                   "cobevt_pack_rgb4_padded")
    return y


def upsample_bilinear2x(x):
    N, H, W, C = x.shape
    y = torch.empty(N, 2 * H, 2 * W, C, dtype=x.dtype, device=x.device)
    with _timed("eltwise", 0, "upsample_bilinear2x"):
        _lib.check(_lib.lib().cobevt_upsample_bilinear2x(_ptr(x), _ptr(y), N, H, W, C, dtype_code(x.dtype), _stream()),
                   "cobevt_upsample_bilinear2x")
    return y


def add_nearest_relu_(y, up):
    """y (N,H,W,C) <- relu(y + nearest_resize(up (N,h,w,C)))  in place."""
    N, H, W, C = y.shape
    _, h, w, _ = up.shape
    with _timed("eltwise", 0, "add_nearest_relu"):
        _lib.check(_lib.lib().cobevt_add_nearest_relu(_ptr(y), _ptr(up), N, H, W, h, w, C, dtype_code(y.dtype), _stream()),
                   "cobevt_add_nearest_relu")
    return y


def invert_matrices(m):
    """(..., d, d) fp32 with d in {3, 4} -> inverse."""
    m = m.float().contiguous()
    d = m.shape[-1]
    out = torch.empty_like(m)
    with _timed("eltwise", 0, "invert_matrices"):
        _lib.check(_lib.lib().cobevt_invert_matrices(_ptr(m), _ptr(out), m.numel() // (d * d), d, _stream()),
                   "cobevt_invert_matrices")
    return out


def maxpool3x3s2(x):
    N, H, W, C = x.shape
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty(N, OH, OW, C, dtype=x.dtype, device=x.device)
    with _timed("eltwise", 0, "maxpool3x3s2_nhwc"):
        _lib.check(_lib.lib().cobevt_maxpool3x3s2_nhwc(_ptr(x), _ptr(y), N, H, W, C, dtype_code(x.dtype), _stream()),
                   "cobevt_maxpool3x3s2_nhwc")
    return y


def bn_relu2(x, sa, ba, sb=None, bb=None):
    C = x.shape[-1]
    rows = x.numel() // C
    ya = torch.empty_like(x)
    yb = torch.empty_like(x) if sb is not None else None
    with _timed("eltwise", 0, "bn_relu2"):
        _lib.check(_lib.lib().cobevt_bn_relu2(_ptr(x), _f32(sa), _f32(ba), _ptr(ya), _ptr(sb), _ptr(bb), _ptr(yb),
                                              rows, C, dtype_code(x.dtype), _stream()), "cobevt_bn_relu2")
    return ya, yb


def ray_embed(I_inv, E_inv, w_img, w_cam, h, w, image_h, image_w, C, dtype, invert_intrinsic=False):
    """I_inv: (NV,3,3) inverse intrinsics, or the intrinsics themselves with invert_intrinsic=True."""
    NV = I_inv.shape[0]
    out = torch.empty(NV, h, w, C, dtype=dtype, device=I_inv.device)
    with _timed("eltwise", 0, "ray_embed"):
        _lib.check(_lib.lib().cobevt_ray_embed(_f32(I_inv), _f32(E_inv), _f32(w_img), _f32(w_cam), _ptr(out), NV, h, w,
                                               float(image_h), float(image_w), C, 1 if invert_intrinsic else 0,
                                               dtype_code(dtype), _stream()),
                   "cobevt_ray_embed")
    return out


def bev_query(x, world, w_bev, b_bev, w_cam, E_inv, BL, n, P, eps=1e-5):
    """x: (rows, C) 16-bit with rows == BL*P or rows == P (broadcast prior)."""
    C = x.shape[-1]
    x_rows = x.numel() // C
    out = torch.empty(BL, n, P, C, dtype=x.dtype, device=x.device)
    with _timed("eltwise", 0, "bev_query"):
        _lib.check(_lib.lib().cobevt_bev_query(_ptr(x), x_rows, _f32(world), _f32(w_bev), _f32(b_bev), _f32(w_cam),
                                               _f32(E_inv), _ptr(out), BL, n, P, C, eps, dtype_code(x.dtype), _stream()),
                   "cobevt_bev_query")
    return out


def layernorm_rows(x, gamma=None, beta=None, eps=1e-5):
    C = x.shape[-1]
    assert x.is_contiguous()
    y = torch.empty_like(x)
    with _timed("eltwise", 0, "layernorm_rows"):
        _lib.check(_lib.lib().cobevt_layernorm_rows(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), x.numel() // C, C, eps,
                                                    dtype_code(x.dtype), _stream()), "cobevt_layernorm_rows")
    return y


def mean_views_ln(x, eps=1e-5):
    """x (B, L, P, C) -> (B, P, C)."""
    B, L, P, C = x.shape
    y = torch.empty(B, P, C, dtype=x.dtype, device=x.device)
    with _timed("eltwise", 0, "mean_views_ln"):
        _lib.check(_lib.lib().cobevt_mean_views_ln(_ptr(x), _ptr(y), B, L, P, C, eps, dtype_code(x.dtype), _stream()),
                   "cobevt_mean_views_ln")
    return y


def sttf_regroup(x, tm, record_len, B, L, discrete_ratio, downsample_rate, use_roi_mask):
    """x (sumL, H, W, C) 16-bit; tm (B, L, 4, 4) fp32; record_len (B,) int64 -- all on device."""
    _, H, W, C = x.shape
    assert tm.dtype == torch.float32 and tm.is_contiguous() and record_len.dtype == torch.int64
    out = torch.empty(B, L, H, W, C, dtype=x.dtype, device=x.device)
    mask = torch.empty(B, H, W, L, dtype=torch.uint8, device=x.device)
    with _timed("eltwise", 0, "sttf_regroup"):
        _lib.check(_lib.lib().cobevt_sttf_regroup(_ptr(x), _ptr(tm), _ptr(record_len), _ptr(out), _ptr(mask), B, L, H, W, C,
                                                  float(discrete_ratio), float(downsample_rate), 1 if use_roi_mask else 0,
                                                  dtype_code(x.dtype), _stream()), "cobevt_sttf_regroup")
    return out, mask


def _any_code(dt):
    return 2 if dt == torch.float32 else dtype_code(dt)


def nchw_to_nhwc(x, dtype, cpad=None):
    """(N, C, H, W) any float dtype -> (N, H, W, Cpad) `dtype`."""
    x = x.contiguous()
    N, C, H, W = x.shape
    cp = cpad or C
    y = torch.empty(N, H, W, cp, dtype=dtype, device=x.device)
    with _timed("eltwise", 0, "permute"):
        _lib.check(_lib.lib().cobevt_permute(_ptr(x), _ptr(y), N, C, H, W, cp, 1, _any_code(x.dtype), _any_code(dtype), _stream()),
                   "cobevt_permute")
    return y


def nhwc_to_nchw(x, dtype, c=None):
    """(N, H, W, Cpad) -> (N, C, H, W) `dtype` (first C channels)."""
    assert x.is_contiguous()
    N, H, W, cp = x.shape
    C = c or cp
    y = torch.empty(N, C, H, W, dtype=dtype, device=x.device)
    with _timed("eltwise", 0, "permute"):
        _lib.check(_lib.lib().cobevt_permute(_ptr(x), _ptr(y), N, C, H, W, cp, 0, _any_code(x.dtype), _any_code(dtype), _stream()),
                   "cobevt_permute")
    return y


def softmax_argmax_nchw(seg):
    """seg (N, C, H, W) fp32 -> (softmax over C (N, C, H, W) fp32, argmax (N, H, W) int64)."""
    seg = seg.float().contiguous()
    N, C, H, W = seg.shape
    prob = torch.empty_like(seg)
    amap = torch.empty(N, H, W, dtype=torch.int64, device=seg.device)
    with _timed("eltwise", 0, "softmax_argmax"):
        _lib.check(_lib.lib().cobevt_softmax_argmax(_ptr(seg), _ptr(prob), _ptr(amap), N, C, H * W, _stream()),
                   "cobevt_softmax_argmax")
    return prob, amap


def resize_nearest(x, scale):
    """F.interpolate(x, scale_factor=scale, recompute_scale_factor=False) (default 'nearest') on NHWC 16-bit."""
    import math
    N, H, W, C = x.shape
    OH, OW = int(math.floor(H * scale)), int(math.floor(W * scale))
    y = torch.empty(N, OH, OW, C, dtype=x.dtype, device=x.device)
    inv = 1.0 / scale
    with _timed("eltwise", 0, "resize_nearest"):
        _lib.check(_lib.lib().cobevt_resize_nearest(_ptr(x), _ptr(y), N, H, W, C, OH, OW, inv, inv, _stream()),
                   "cobevt_resize_nearest")
    return y


def dwconv_swish(x, w, bias, K, stride, pad_top, pad_left, OH, OW, sums=None):
    """x (N,H,W,C) 16-bit; w (K,K,C), bias (C) fp32 (BatchNorm folded) -> swish(dwconv) (N,OH,OW,C); sums (N,C) fp32
    accumulates the per-channel output sums (squeeze)."""
    N, H, W, C = x.shape
    y = torch.empty(N, OH, OW, C, dtype=x.dtype, device=x.device)
    with _timed("eltwise", 2 * N * OH * OW * C * K * K, f"dwconv_swish k{K} s{stride} C={C} {H}x{W}"):
        _lib.check(_lib.lib().cobevt_dwconv_swish(_ptr(x), _f32(w), _f32(bias), _ptr(y), _ptr(sums), N, H, W, C, K, stride,
                                                  pad_top, pad_left, OH, OW, dtype_code(x.dtype), _stream()),
                   "cobevt_dwconv_swish")
    return y


def se_gate(sums, hw, w1, b1, w2, b2):
    N, C = sums.shape
    gate = torch.empty_like(sums)
    with _timed("eltwise", 0, "se_gate"):
        _lib.check(_lib.lib().cobevt_se_gate(_f32(sums), 1.0 / hw, _f32(w1), _f32(b1), _f32(w2), _f32(b2), _ptr(gate), N, C,
                                             w1.shape[0], _stream()), "cobevt_se_gate")
    return gate


def scale_channels_(x, gate):
    """x (N,H,W,C) 16-bit *= gate (N,C) fp32, in place."""
    N, H, W, C = x.shape
    with _timed("eltwise", 0, "scale_channels"):
        _lib.check(_lib.lib().cobevt_scale_channels(_ptr(x), _f32(gate), _ptr(x), N, H * W, C, dtype_code(x.dtype), _stream()),
                   "cobevt_scale_channels")
    return x


def bev_q_analytic(x, pc_wx, tab, consts, world, E_inv, BL, n, P, eps=1e-5):
    """Stage-0 query projection without the LayerNorm round trip: x (rows, C) 16-bit with rows == P (shared prior) or
    BL*P; pc_wx = PackedConv(W', no bias); tab (C,16) fp32 channel table, consts = (qa, qb, qd, ma, mb) (see
    include/cobevt_b200.h).  Returns q (BL, n, P, C) 16-bit."""
    C = x.shape[-1]
    rows = x.numel() // C
    qx = linear_tokens(x.reshape(rows, C), pc_wx, want16=False, out32=True)[1]
    stats = torch.empty(rows, 12, dtype=torch.float32, device=x.device)
    with _timed("eltwise", 0, "row_stats9"):
        _lib.check(_lib.lib().cobevt_row_stats9(_ptr(x), _f32(tab), _ptr(stats), rows, C, dtype_code(x.dtype), _stream()),
                   "cobevt_row_stats9")
    out = torch.empty(BL, n, P, C, dtype=x.dtype, device=x.device)
    qa, qb, qd, ma, mb = consts
    with _timed("eltwise", 0, "bev_q_analytic"):
        _lib.check(_lib.lib().cobevt_bev_q_analytic(_f32(qx), rows, _f32(stats), _f32(world), _f32(tab), _f32(E_inv), _ptr(out),
                                                    BL, n, P, C, qa, qb, qd, ma, mb, eps, dtype_code(x.dtype), _stream()),
                   "cobevt_bev_q_analytic")
    return out


def se_scale_(x, sums, w1, b1, w2, b2):
    """squeeze-excite in one launch: x (N,H,W,C) 16-bit *= sigmoid(W2 swish(W1 mean + b1) + b2), mean = sums / (H W)."""
    N, H, W, C = x.shape
    with _timed("eltwise", 0, "se_scale"):
        _lib.check(_lib.lib().cobevt_se_scale(_ptr(x), _f32(sums), 1.0 / (H * W), _f32(w1), _f32(b1), _f32(w2), _f32(b2), N, H * W, C,
                                              w1.shape[0], dtype_code(x.dtype), _stream()), "cobevt_se_scale")
    return x
