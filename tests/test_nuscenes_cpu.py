"""nuScenes SinBEVT (CVT fork): oracle pinned to the unmodified reference (CPU, 1e-5) and drop-in
state_dict parity of cobevt_b200.nuscenes.  Needs /root/reference (skipped elsewhere)."""
import copy

import pytest
import torch

import refimport
import nusc_common as nc

pytestmark = pytest.mark.reference


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def ref():
    return refimport.load_nuscenes()


@pytest.mark.parametrize("case", ["native200_1cam", "c2_bev256_2cam"])
def test_cvt_oracle_vs_reference(ref, case):
    from cobevt_b200 import configs, synth
    if case == "native200_1cam":       # BASELINE C1 secondary: 1 cam 224x224, BEV 200, live key padding
        H, W, n = 224, 224, 1
        cfg = configs.nuscenes_cvt(H, W, 200, feat_win=((6, 6), (6, 6), (14, 14)))
    else:                              # C2 geometry (BEV 256, 7x15 windows), 2 cameras to keep CPU time low
        H, W, n = 224, 480, 2
        cfg = configs.nuscenes_c2()
    m = nc.build_reference(ref, cfg, H, W)
    sd = synth.synth_state_dict(m, seed=0)
    m.load_state_dict(sd)
    batch = synth.synth_nuscenes_batch(H, W, b=1, n_cams=n, seed=2)
    with torch.no_grad():
        want = m(copy.deepcopy(batch))
        got = nc.oracle_forward(batch, sd, cfg)
    for k in want:
        assert got[k].shape == want[k].shape
        assert rel(got[k], want[k]) < 1e-5, (k, rel(got[k], want[k]))


def test_nuscenes_state_dict_is_drop_in(ref):
    from cobevt_b200 import configs
    cfg = configs.nuscenes_cvt()
    r = nc.build_reference(ref, cfg, 224, 480).state_dict()
    o = nc.build_ours(cfg, 224, 480).state_dict()
    # the reference-side adaptor wraps a full torchvision resnet (layer4 + fc exist on both sides)
    assert sorted(r.keys()) == sorted(o.keys())
    for k in r:
        assert r[k].shape == o[k].shape, k


_FakeEffNetFeatures = nc.FakeEffNetFeatures
_load_shipped_ckpt = nc.load_shipped_ckpt


def test_real_weights_shipped_checkpoint(ref):
    """The trained SinBEVT weights shipped with the reference (nuscenes/logs/*.ckpt): FAX stages, Bottlenecks,
    downsample blocks, decoder and logits head -- reference vs oracle at 1e-5 on the native 200x200 config."""
    from cobevt_b200 import configs, synth
    H, W = 224, 480
    cfg = configs.nuscenes_cvt()
    e = copy.deepcopy(cfg['encoder'])
    enc = ref.epa.PyramidAxialEncoder(_FakeEffNetFeatures(H, W), e['cross_view'], e['cross_view_swap'],
                                      e['bev_embedding'], e['self_attn'], e['dim'], e['middle'], e['scale'])
    m = ref.cvt.CrossViewTransformer(enc, ref.decoder.Decoder(128, [128, 128, 64], True, 2), 64,
                                     {'bev': [0, 1], 'center': [1, 2]}).eval()
    ck = {k[len("backbone."):]: v for k, v in _load_shipped_ckpt().items() if k.startswith("backbone.")}
    ck = {k: v for k, v in ck.items() if not k.startswith("encoder.backbone.")}
    missing, unexpected = m.load_state_dict(ck, strict=False)
    assert not unexpected and all(k.startswith("encoder.backbone") for k in missing), (missing[:3], unexpected[:3])
    sd = {k: v.float() for k, v in m.state_dict().items()}
    batch = synth.synth_nuscenes_batch(H, W, b=1, n_cams=6, seed=3)
    from oracle import cobevt_oracle as orc
    with torch.no_grad():
        want = m(copy.deepcopy(batch))
        got = orc.cross_view_transformer(batch, sd, cfg, lambda img: _FakeEffNetFeatures(H, W)(img))
    # trained weights have a much larger dynamic range than the synthetic ones: fp32 re-association noise
    # (different summation order in the attn2 de-duplication / folded reductions) reaches 1.3e-5 here
    for k in want:
        assert rel(got[k], want[k]) < 5e-5, (k, rel(got[k], want[k]))
