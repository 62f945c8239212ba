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
