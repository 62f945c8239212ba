"""ctypes binding of libcobevt_sm100.so (the C ABI declared in include/cobevt_b200.h).

There is no fallback: if the shared library is missing or a call fails the error is
raised.  The library is built in-tree by ``cobevt_b200/csrc/build.sh`` (see
``__graft_entry__.build``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# COBEVT_LIB_PATH: load another build of the SAME library (kernel A/B timing); never a different backend
LIB_PATH = os.environ.get("COBEVT_LIB_PATH") or os.path.join(_HERE, "libcobevt_sm100.so")

c_i32 = ctypes.c_int32
c_f32 = ctypes.c_float
c_vp = ctypes.c_void_p


class ConvParams(ctypes.Structure):
    _fields_ = [
        ("x", c_vp), ("w", c_vp), ("bias", c_vp), ("res", c_vp), ("res32", c_vp),
        ("y", c_vp), ("y32", c_vp), ("y_ln", c_vp), ("ln_gamma", c_vp), ("ln_beta", c_vp),
        ("N", c_i32), ("Hin", c_i32), ("Win", c_i32), ("Cin", c_i32),
        ("OH", c_i32), ("OW", c_i32), ("Cout", c_i32),
        ("KH", c_i32), ("KW", c_i32),
        ("stride_h", c_i32), ("stride_w", c_i32), ("pad_h", c_i32), ("pad_w", c_i32), ("res_rows", c_i32),
        ("in_upsample", c_i32), ("act", c_i32), ("dtype", c_i32), ("block_n", c_i32),
        ("ln_eps", c_f32),
    ]


class MlpParams(ctypes.Structure):
    _fields_ = [
        ("x_ln", c_vp), ("w1", c_vp), ("b1", c_vp), ("w2", c_vp), ("b2", c_vp), ("res", c_vp), ("y", c_vp), ("y_ln", c_vp),
        ("ln_gamma", c_vp), ("ln_beta", c_vp),
        ("M", c_i32), ("C", c_i32), ("H", c_i32), ("dtype", c_i32), ("ln_eps", c_f32),
    ]


class TokenMap(ctypes.Structure):
    _fields_ = [
        ("views", c_i32), ("H", c_i32), ("W", c_i32), ("Hv", c_i32), ("Wv", c_i32),
        ("w1", c_i32), ("w2", c_i32), ("grid", c_i32),
    ]


class AttnParams(ctypes.Structure):
    _fields_ = [
        ("q", c_vp), ("k", c_vp), ("v", c_vp), ("out", c_vp),
        ("k_pad", c_vp), ("v_pad", c_vp), ("bias", c_vp), ("key_mask", c_vp),
        ("qmap", TokenMap), ("kvmap", TokenMap),
        ("B", c_i32), ("heads", c_i32),
        ("q_stride", c_i32), ("kv_stride", c_i32), ("out_stride", c_i32), ("q_mean", c_i32), ("scale", c_f32), ("dtype", c_i32),
        ("part", c_vp), ("kv_view0", c_i32), ("kv_views_total", c_i32), ("kv_segments", c_i32),
        ("part_segment_stride", ctypes.c_int64),
        ("bias_layout", c_i32),
    ]


_lib = None


def lib():
    """Load (once) and return the shared library; raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with cobevt_b200/csrc/build.sh "
                "(or __graft_entry__.build()); there is no fallback path")
        L = ctypes.CDLL(LIB_PATH)
        L.cobevt_version.restype = ctypes.c_int
        L.cobevt_last_error.restype = ctypes.c_char_p
        L.cobevt_conv_packed_bytes.restype = ctypes.c_int64
        L.cobevt_conv_packed_bytes.argtypes = [c_i32, c_i32, c_i32]
        L.cobevt_conv_pack_weight.restype = ctypes.c_int
        L.cobevt_conv_pack_weight.argtypes = [c_vp, c_i32, c_i32, c_i32, c_vp]
        L.cobevt_conv2d_nhwc_fwd.restype = ctypes.c_int
        L.cobevt_conv2d_nhwc_fwd.argtypes = [ctypes.POINTER(ConvParams), c_vp]
        L.cobevt_token_mlp_fwd.restype = ctypes.c_int
        L.cobevt_token_mlp_fwd.argtypes = [ctypes.POINTER(MlpParams), c_vp]
        L.cobevt_window_attn_fwd.restype = ctypes.c_int
        L.cobevt_window_attn_fwd.argtypes = [ctypes.POINTER(AttnParams), c_vp]
        L.cobevt_attn_workspace_bytes.restype = ctypes.c_int64
        L.cobevt_attn_workspace_bytes.argtypes = [ctypes.POINTER(AttnParams)]
        L.cobevt_attn_resident_keys.restype = ctypes.c_int32
        L.cobevt_attn_resident_keys.argtypes = []
        L.cobevt_attn_merge.restype = ctypes.c_int
        L.cobevt_attn_merge.argtypes = [c_vp, c_i32, ctypes.c_int64, c_vp, ctypes.POINTER(TokenMap), c_i32, c_i32, c_i32,
                                        c_i32, c_i32, c_vp]
        for name, argt in _ELTWISE_SIGS.items():
            fn = getattr(L, name)
            fn.restype = ctypes.c_int
            fn.argtypes = argt
        # the ctypes declarations above must describe the structs exactly as the library was compiled
        L.cobevt_struct_bytes.restype = ctypes.c_int32
        L.cobevt_struct_bytes.argtypes = [c_i32]
        for which, cls in enumerate((ConvParams, AttnParams, MlpParams, TokenMap)):
            want = int(L.cobevt_struct_bytes(which))
            if want != ctypes.sizeof(cls):
                raise RuntimeError(f"{LIB_PATH}: sizeof({cls.__name__}) is {want} in the library but {ctypes.sizeof(cls)} "
                                   "in cobevt_b200/_lib.py -- rebuild the library (cobevt_b200/csrc/build.sh)")
        _lib = L
    return _lib


# element-wise / data-movement entry points: name -> argtypes (see include/cobevt_b200.h)
c_i64 = ctypes.c_int64
_ELTWISE_SIGS = {
    "cobevt_pack_rgb8": [c_vp, c_vp, c_i64, c_i32, c_vp],
    "cobevt_pack_rgb4_padded": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp],
    "cobevt_upsample_bilinear2x": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "cobevt_add_nearest_relu": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "cobevt_invert_matrices": [c_vp, c_vp, c_i32, c_i32, c_vp],
    "cobevt_maxpool3x3s2_nhwc": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "cobevt_bn_relu2": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp],
    "cobevt_ray_embed": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_f32, c_f32, c_i32, c_i32, c_i32, c_vp],
    "cobevt_bev_query": [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_i32, c_vp],
    "cobevt_layernorm_rows": [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_f32, c_i32, c_vp],
    "cobevt_mean_views_ln": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_i32, c_vp],
    "cobevt_sttf_regroup": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_f32, c_f32, c_i32, c_i32, c_vp],
    "cobevt_permute": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "cobevt_stage_rgb_u8": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp],
    "cobevt_softmax_argmax": [c_vp, c_vp, c_vp, c_i32, c_i32, c_i64, c_vp],
    "cobevt_resize_nearest": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_f32, c_f32, c_vp],
    "cobevt_dwconv_swish": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32,
                            c_i32, c_i32, c_vp],
    "cobevt_se_gate": [c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp],
    "cobevt_scale_channels": [c_vp, c_vp, c_vp, c_i32, c_i64, c_i32, c_i32, c_vp],
    "cobevt_se_scale": [c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_i32, c_i64, c_i32, c_i32, c_i32, c_vp],
    "cobevt_row_stats9": [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp],
    "cobevt_bev_q_analytic": [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_f32, c_f32,
                              c_f32, c_f32, c_f32, c_i32, c_vp],
}


def check(rc, what):
    if rc != 0:
        msg = lib().cobevt_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


EXPORTED_SYMBOLS = [
    "cobevt_version", "cobevt_last_error", "cobevt_struct_bytes", "cobevt_conv_packed_bytes",
    "cobevt_conv_pack_weight", "cobevt_conv2d_nhwc_fwd", "cobevt_token_mlp_fwd", "cobevt_window_attn_fwd", "cobevt_attn_merge", "cobevt_attn_resident_keys", "cobevt_attn_workspace_bytes",
] + list(_ELTWISE_SIGS)
