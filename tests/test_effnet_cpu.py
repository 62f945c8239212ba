"""EfficientNet-B4 prefix: structure of the mirror module and of the oracle restatement.
The arithmetic belongs to efficientnet_pytorch==0.7.1, absent here: numerics are 'parity unpinned'; what IS pinned is
that the shipped checkpoint's 236 backbone tensors strict-load into the mirror (names + shapes), that the oracle
consumes exactly those tensors, and the feature shapes the reference documents (efficientnet.py:34)."""
import pytest
import torch


def test_effnet_shapes_and_oracle_runs():
    from cobevt_b200 import synth, nuscenes as ns
    from oracle import cobevt_oracle as orc
    ext = ns.EfficientNetExtractor(['reduction_2', 'reduction_3', 'reduction_4'], 224, 480).eval()
    assert [tuple(s) for s in ext.output_shapes] == [(1, 32, 56, 120), (1, 56, 28, 60), (1, 112, 14, 30)]
    assert sum(1 for _ in ext.state_dict()) == 236
    sd = synth.synth_state_dict(ext, seed=0)
    with torch.no_grad():
        f = orc.efficientnet_b4_prefix(torch.rand(1, 3, 224, 480), sd, "")
    assert [tuple(t.shape) for t in f] == [(1, 32, 56, 120), (1, 56, 28, 60), (1, 112, 14, 30)]
    # static 'same' padding follows the nominal 380-pixel chain: (0,1) for the stride-2 3x3s, (2,2) for every 5x5
    assert orc.static_same_pad(380, 3, 2) == (0, 1) and orc.static_same_pad(190, 3, 2) == (0, 1)
    assert orc.static_same_pad(95, 5, 2) == (2, 2) and orc.static_same_pad(48, 3, 2) == (0, 1)


@pytest.mark.reference
def test_shipped_checkpoint_backbone_strict_loads():
    import nusc_common as nc
    from cobevt_b200 import nuscenes as ns
    from oracle import cobevt_oracle as orc
    ck = nc.load_shipped_ckpt()
    bb = {k[len("backbone.encoder.backbone."):]: v for k, v in ck.items() if k.startswith("backbone.encoder.backbone.")}
    assert len(bb) == 236
    ext = ns.EfficientNetExtractor(['reduction_2', 'reduction_3', 'reduction_4'], 224, 480).eval()
    ext.load_state_dict(bb, strict=True)
    with torch.no_grad():
        f = orc.efficientnet_b4_prefix(torch.rand(1, 3, 224, 480), {k: v.float() for k, v in ext.state_dict().items()}, "")
    assert all(torch.isfinite(t).all() for t in f)
    # the FAX stages of the same checkpoint expect exactly these channel counts (cross_views.i.feature_linear.0)
    for i, c in enumerate((32, 56, 112)):
        assert ck[f"backbone.encoder.cross_views.{i}.feature_linear.0.weight"].shape == (c,)
        assert f[i].shape[1] == c
