"""CPU: the C-ABI shared library loads and exports every symbol include/cobevt_b200.h declares
(no compute calls -- there is no GPU here), and host-side argument validation works."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cobevt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cobevt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cobevt_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names


def test_version_and_error_string():
    from cobevt_b200 import _lib
    L = _lib.lib()
    assert L.cobevt_version() >= 100
    p = _lib.ConvParams()            # all-null -> shape error, message set, no CUDA touched
    rc = L.cobevt_conv2d_nhwc_fwd(ctypes.byref(p), None)
    assert rc == -1
    assert b"null" in L.cobevt_last_error()
    a = _lib.AttnParams()
    assert L.cobevt_window_attn_fwd(ctypes.byref(a), None) == -1


def test_struct_layouts_match_the_library():
    """ctypes declarations vs the structs as compiled (cobevt_struct_bytes); field offsets the kernels rely on follow from
    the declaration order, so the sizes pin accidental drift (the loader refuses a mismatching library as well)."""
    from cobevt_b200 import _lib
    L = _lib.lib()
    for which, cls in enumerate((_lib.ConvParams, _lib.AttnParams, _lib.MlpParams, _lib.TokenMap)):
        assert L.cobevt_struct_bytes(which) == ctypes.sizeof(cls), cls.__name__
    assert L.cobevt_struct_bytes(99) == -1
    assert _lib.AttnParams.bias_layout.offset + 4 <= ctypes.sizeof(_lib.AttnParams)
    a = _lib.AttnParams()
    a.bias_layout = 7                                           # rejected on the host before anything else is read
    a.q = a.k = a.v = a.out = ctypes.c_void_p(256)
    assert L.cobevt_window_attn_fwd(ctypes.byref(a), None) == -1


def test_attn_bias_key_group_major_copy():
    """ops.AttnBias: kq[h, k/4, q, :] == raw[h, q, 4*(k/4) : 4*(k/4)+4] (the layout bias_layout = 1 addresses)."""
    import torch
    from cobevt_b200 import ops
    raw = torch.arange(2 * 6 * 8, dtype=torch.float32).reshape(2, 6, 8)
    b = ops.AttnBias(raw)
    assert b.kq.shape == (2, 2, 6, 4) and b.kq.is_contiguous()
    for h in range(2):
        for q in range(6):
            for g in range(2):
                assert torch.equal(b.kq[h, g, q], raw[h, q, 4 * g:4 * g + 4])
    assert ops.AttnBias(torch.zeros(1, 3, 6)).kq is None       # key count not a multiple of 4: row-major only


def test_attention_host_validation_and_capacity():
    """Host-side checks of the attention entry points (no kernel is launched: every call fails validation)."""
    from cobevt_b200 import _lib
    L = _lib.lib()
    cap = L.cobevt_attn_resident_keys()
    assert cap > 0 and cap % 64 == 0 and cap >= 630          # the north_star windows (630 keys) must stay resident
    a = _lib.AttnParams()
    a.kv_segments = 3                                          # segments without a partial buffer / pointers
    assert L.cobevt_window_attn_fwd(ctypes.byref(a), None) == -1
    m = _lib.TokenMap()
    assert L.cobevt_attn_merge(None, 2, 0, None, ctypes.byref(m), 1, 4, 0, 0, 0, None) == -1
    assert b"attn_merge" in L.cobevt_last_error()


def test_weight_packing_layout():
    """Packed image = [n_tile][k_block][row][64] with the SWIZZLE_128B chunk XOR."""
    import torch
    from cobevt_b200 import ops
    cout, cin = 40, 72                      # ragged: 1 n-tile of 64, 2 k-blocks (second half empty)
    w = torch.arange(cout * cin, dtype=torch.float32).reshape(cout, cin) % 1000
    pc = ops.PackedConv(w, None, "cpu", dtype=torch.float16, block_n=64)
    img = pc.w.view(torch.float16).reshape(1, 2, 64, 64)       # nt kb row 64
    for (r, k) in [(0, 0), (3, 9), (39, 71), (17, 63), (5, 64)]:
        kb, kk = divmod(k, 64)
        chunk = (kk // 8) ^ (r % 8)
        assert img[0, kb, r, chunk * 8 + kk % 8].item() == w[r, k].item()
    assert img[0, :, 40:].abs().sum().item() == 0             # rows past Cout are zero
    assert img[0, 1, :, :].reshape(64, 8, 8)[0].abs().sum().item() > 0


def test_modules_refuse_cpu_and_training():
    import torch
    import cobevt_b200 as cb
    from cobevt_b200 import configs
    m = cb.SwapFusionEncoder(configs.small()['fax_fusion'])
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 5, 128, 8, 8))                        # training mode
    m.eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 5, 128, 8, 8), torch.ones(1, 8, 8, 1, 5))   # CPU tensor: no fallback
