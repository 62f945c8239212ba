"""GPU parity of the drop-in modules (CUDA kernels through the C ABI) against the CPU oracle
on identical seeded weights and inputs.

Tolerance: the kernels store activations as fp16 (fp32 accumulation); the bar is
max|err| <= tol * max|ref|.  Whole-model bound 4.5e-3 = this repo's measured error (1.9e-3 .. 2.9e-3) x 1.5,
below what the unmodified reference itself moves under fp16 autocast (2.3e-3 .. 5.4e-3,
tests/golden/autocast_error.json); see tests/test_golden.py.
"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _build(cfg, cls_name="CorpBEVT", seed=0):
    import cobevt_b200 as cb
    from cobevt_b200 import synth
    m = getattr(cb, cls_name)(copy.deepcopy(cfg)).eval()
    sd = synth.synth_state_dict(m, seed=seed)
    m.load_state_dict(sd)
    return m, sd


@pytest.mark.parametrize("record_len,n_cams", [((3, 2), 2), ((5,), 4), ((1, 4, 2), 1)])
def test_corpbevt_small_vs_oracle(record_len, n_cams):
    from cobevt_b200 import configs, synth
    from oracle import cobevt_oracle as orc
    cfg = configs.small()
    m, sd = _build(cfg)
    batch = synth.synth_batch(cfg, record_len=record_len, n_cams=n_cams, seed=1)
    with torch.no_grad():
        ref = orc.corpbevt(copy.deepcopy(batch), sd, cfg)
        out = m.cuda()(synth.to_device(batch, "cuda"))
    torch.cuda.synchronize()
    for k in ("dynamic_seg", "static_seg"):
        assert out[k].shape == ref[k].shape
    e = relerr(out["dynamic_seg"], ref["dynamic_seg"])
    print("corpbevt small rel err", e)
    assert e < 4.5e-3, e
    assert out["static_seg"].abs().max().item() == 0.0


def test_fax_fused_small_vs_oracle():
    from cobevt_b200 import configs, synth
    from oracle import cobevt_oracle as orc
    cfg = configs.fax_fused_args(configs.small())
    m, sd = _build(cfg, "FaxFusedTransformer")
    batch = synth.synth_batch(dict(cfg, max_cav=1), record_len=(1, 1), n_cams=3, seed=2)
    batch["inputs"] = batch["inputs"].reshape(2, 1, *batch["inputs"].shape[2:])
    with torch.no_grad():
        ref = orc.fax_fused_transformer(copy.deepcopy(batch), sd, cfg)
        out = m.cuda()(synth.to_device(batch, "cuda"))
    e = relerr(out["dynamic_seg"], ref["dynamic_seg"])
    print("fax fused small rel err", e)
    assert e < 4.5e-3, e


def test_full_size_batch_independence():
    """BASELINE full size (5 agents x 6 cams x 224x480 -> 256x256 BEV): a sample's output must not
    depend on what else is in the batch (agent rows are independent until FuseBEVT, which only mixes
    agents of the same sample).  Bit-exact: the kernels are deterministic and row-independent."""
    from cobevt_b200 import configs, synth
    cfg = configs.north_star()
    m, _ = _build(cfg)
    m = m.cuda()
    two = synth.synth_batch(cfg, record_len=(5, 3), n_cams=6, seed=11)
    one = {k: (v[:5] if k in ("inputs", "intrinsic", "extrinsic") else v[:1]) for k, v in two.items()}
    with torch.no_grad():
        o2 = m(synth.to_device(two, "cuda"))["dynamic_seg"]
        o1 = m(synth.to_device(one, "cuda"))["dynamic_seg"]
    torch.cuda.synchronize()
    assert o2.shape == (2, 1, 2, 256, 256)
    assert torch.isfinite(o2).all()
    assert torch.equal(o2[:1], o1)


def test_sharded_graph_replay_equals_module_forward():
    """parallel.ShardedCorpBEVT (world 1, CUDA-graph replay) launches the same kernels as
    CorpBEVT.forward: outputs are bit-identical, also on the second (replayed) call."""
    from cobevt_b200 import configs, synth
    from cobevt_b200.parallel import ShardedCorpBEVT
    cfg = configs.small()
    m, _ = _build(cfg)
    m = m.cuda()
    batch = synth.to_device(synth.synth_batch(cfg, record_len=(3, 2), n_cams=2, seed=1), "cuda")
    sh = ShardedCorpBEVT(m, 0, 1, graphs=True)
    with torch.no_grad():
        want = m(dict(batch))["dynamic_seg"].clone()
        for _ in range(3):
            got = sh(batch, [3, 2], batch["transformation_matrix"])["dynamic_seg"]
            torch.cuda.synchronize()
            assert torch.equal(got, want)
