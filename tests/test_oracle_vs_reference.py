"""Pins the CPU oracle (oracle/cobevt_oracle.py) to the UNMODIFIED reference imported from
/root/reference (authoring container only; skipped elsewhere).  fp32, tolerance 1e-5 relative
 -- these are the C1 'plumbing + numerics gate' runs."""
import copy

import pytest
import torch

import refimport

pytestmark = pytest.mark.reference


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def ref():
    return refimport.load_opv2v()


def _ref_model(ref, cfg, kind="corpbevt"):
    from cobevt_b200 import synth
    cls = ref.corpbevt.CorpBEVT if kind == "corpbevt" else ref.fax_fused.FaxFusedTransformer
    m = cls(copy.deepcopy(cfg)).eval()
    sd = synth.synth_state_dict(m, seed=0)
    m.load_state_dict(sd)
    return m, sd


@pytest.mark.parametrize("record_len,n_cams", [((3, 2), 2), ((5,), 4), ((1, 4, 2), 1)])
def test_corpbevt_small(ref, record_len, n_cams):
    from cobevt_b200 import configs, synth
    from oracle import cobevt_oracle as orc
    cfg = configs.small()
    m, sd = _ref_model(ref, cfg)
    batch = synth.synth_batch(cfg, record_len=record_len, n_cams=n_cams, seed=1)
    with torch.no_grad():
        want = m(copy.deepcopy(batch))
        got = orc.corpbevt(copy.deepcopy(batch), sd, cfg)
    assert rel(got["dynamic_seg"], want["dynamic_seg"]) < 1e-5
    assert torch.equal(got["static_seg"], want["static_seg"])


def test_corpbevt_no_mask_variant(ref):
    from cobevt_b200 import configs, synth
    from oracle import cobevt_oracle as orc
    cfg = configs.small()
    cfg["fax_fusion"]["mask"] = False
    m, sd = _ref_model(ref, cfg)
    batch = synth.synth_batch(cfg, record_len=(2, 5), n_cams=2, seed=3)
    with torch.no_grad():
        want = m(copy.deepcopy(batch))
        got = orc.corpbevt(copy.deepcopy(batch), sd, cfg)
    assert rel(got["dynamic_seg"], want["dynamic_seg"]) < 1e-5


def test_c1_fax_fused_224(ref):
    """BASELINE config C1: single agent, single camera 224x224 SinBEVT forward."""
    from cobevt_b200 import configs, synth
    from oracle import cobevt_oracle as orc
    cfg = configs.fax_fused_args(configs.corpbevt_args(
        image_h=224, image_w=224, bev=224, q_win=((16, 16), (8, 8), (28, 28)),
        feat_win=((4, 4), (2, 2), (7, 7)), self_attn_window=28))
    m, sd = _ref_model(ref, cfg, "fax")
    batch = synth.synth_batch(dict(cfg, max_cav=1), record_len=(1,), n_cams=1, seed=4)
    with torch.no_grad():
        want = m(copy.deepcopy(batch))
        got = orc.fax_fused_transformer(copy.deepcopy(batch), sd, cfg)
    assert want["dynamic_seg"].shape == (1, 1, 2, 224, 224)
    assert rel(got["dynamic_seg"], want["dynamic_seg"]) < 1e-5


def test_north_star_geometry_one_agent(ref):
    """6 cams x 224x480 -> BEV 256 with 7x15 feature windows (630 keys per window)."""
    from cobevt_b200 import configs, synth
    from oracle import cobevt_oracle as orc
    cfg = configs.north_star()
    m, sd = _ref_model(ref, cfg)
    batch = synth.synth_batch(cfg, record_len=(1,), n_cams=6, seed=5)
    with torch.no_grad():
        want = m(copy.deepcopy(batch))
        got = orc.corpbevt(copy.deepcopy(batch), sd, cfg)
    assert rel(got["dynamic_seg"], want["dynamic_seg"]) < 1e-5


def test_stage_with_live_padding(ref):
    """One CrossViewSwapAttention with zero-padded key/value maps (56->60-style), reference vs oracle."""
    from cobevt_b200 import synth
    from oracle import cobevt_oracle as orc
    fax = ref.fax
    torch.manual_seed(0)
    bev = fax.BEVEmbedding(64, 1.0, 40, 40, 100, 100, 0.0, [2, 4, 8])
    cva = fax.CrossViewSwapAttention(feat_height=10, feat_width=20, feat_dim=24, dim=64, index=0,
                                     image_height=80, image_width=160, qkv_bias=True,
                                     q_win_size=[[10, 10]], feat_win_size=[[6, 12]], heads=[2], dim_head=[32],
                                     bev_embedding_flag=[True]).eval()
    sd = synth.synth_state_dict(cva, seed=9)
    cva.load_state_dict(sd)
    g = torch.Generator().manual_seed(1)
    b, n = 2, 3
    x = torch.randn(b, 64, 20, 20, generator=g)
    feat = torch.randn(b, n, 24, 10, 20, generator=g)
    batch = synth.synth_batch({'encoder': {'image_height': 80, 'image_width': 160}, 'max_cav': 1}, (1, 1), n, seed=2)
    I_inv = batch['intrinsic'].reshape(b, n, 3, 3).inverse()
    E_inv = batch['extrinsic'].reshape(b, n, 4, 4)
    with torch.no_grad():
        want = cva(0, x, bev, feat, I_inv, E_inv)
        got = orc.cross_view_swap_attention(x, bev.grid0[:2], feat, I_inv, E_inv, {"cv." + k: v for k, v in sd.items()},
                                            "cv", image_h=80, image_w=160, q_win=(10, 10), feat_win=(6, 12),
                                            heads=2, dim_head=32, bev_embed_flag=True)
    assert rel(got, want) < 1e-5


def test_state_dict_is_drop_in(ref):
    """Our modules expose exactly the reference's state_dict keys and shapes (native corpbevt.yaml)."""
    import cobevt_b200 as cb
    from cobevt_b200 import configs
    cfg = configs.native_opv2v()
    r = ref.corpbevt.CorpBEVT(copy.deepcopy(cfg)).state_dict()
    o = cb.CorpBEVT(copy.deepcopy(cfg)).state_dict()
    assert list(r.keys()) == list(o.keys())
    for k in r:
        assert r[k].shape == o[k].shape, k
        if not r[k].is_floating_point() and "num_batches_tracked" not in k:
            assert torch.equal(r[k], o[k]), k
