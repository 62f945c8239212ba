"""Generates tests/golden/*.pt from the UNMODIFIED reference (run in the authoring container,
where /root/reference exists):   python tests/golden/make_golden.py

Each fixture holds only the seeded-input recipe and the REFERENCE's output; weights and inputs
are regenerated bit-identically from cobevt_b200.synth on any box, so the GPU box (which has no
/root/reference) can compare the CUDA path and the oracle against what the reference produced.
"""
import copy
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refimport  # noqa: E402
from cobevt_b200 import configs, synth  # noqa: E402

CASES = {
    # name: (config builder, model kind, record_len, n_cams, weight seed, input seed)
    "corpbevt_small_ragged": (configs.small, "corpbevt", (3, 2), 2, 0, 1),
    "corpbevt_small_5x4": (configs.small, "corpbevt", (5,), 4, 0, 1),
    "corpbevt_small_nomask": (lambda: _nomask(configs.small()), "corpbevt", (2, 5), 2, 0, 3),
    "faxfused_c1_224": (lambda: configs.fax_fused_args(configs.corpbevt_args(
        image_h=224, image_w=224, bev=224, q_win=((16, 16), (8, 8), (28, 28)), feat_win=((4, 4), (2, 2), (7, 7)),
        self_attn_window=28)), "fax", (1,), 1, 0, 4),
    "corpbevt_northstar_1agent": (configs.north_star, "corpbevt", (1,), 6, 0, 5),
    "corpbevt_northstar_2agents": (configs.north_star, "corpbevt", (2,), 6, 0, 6),
}


NUSC_CASES = {
    # name: (config builder, H, W, n_cams, weight seed, input seed)
    "cvt_native200_1cam": (lambda: configs.nuscenes_cvt(224, 224, 200, feat_win=((6, 6), (6, 6), (14, 14))), 224, 224, 1, 0, 2),
    "cvt_c2_bev256_2cam": (configs.nuscenes_c2, 224, 480, 2, 0, 2),
}


def _nomask(cfg):
    cfg["fax_fusion"]["mask"] = False
    return cfg


def main():
    ref = refimport.load_opv2v()
    for name, (mk, kind, record_len, n_cams, wseed, iseed) in CASES.items():
        cfg = mk()
        cls = ref.corpbevt.CorpBEVT if kind == "corpbevt" else ref.fax_fused.FaxFusedTransformer
        m = cls(copy.deepcopy(cfg)).eval()
        sd = synth.synth_state_dict(m, seed=wseed)
        m.load_state_dict(sd)
        batch = synth.synth_batch(dict(cfg, max_cav=cfg.get("max_cav", 1)), record_len=record_len, n_cams=n_cams, seed=iseed)
        if kind == "fax":
            batch["inputs"] = batch["inputs"].reshape(len(record_len), 1, *batch["inputs"].shape[2:])
        with torch.no_grad():
            out = m(copy.deepcopy(batch))
        dyn = out["dynamic_seg"]
        # fp16 storage keeps the fixtures small; the stored value is exactly what tests compare to
        torch.save({"case": name, "kind": kind, "config": cfg, "record_len": record_len, "n_cams": n_cams,
                    "weight_seed": wseed, "input_seed": iseed, "dynamic_seg": dyn.clone(),
                    "weights_checksum": float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point())),
                    "torch": torch.__version__}, os.path.join(HERE, name + ".pt"))
        print(name, tuple(dyn.shape), float(dyn.abs().max()))
    import nusc_common as nc
    nref = refimport.load_nuscenes()
    for name, (mk, H, W, n_cams, wseed, iseed) in NUSC_CASES.items():
        cfg = mk()
        m = nc.build_reference(nref, cfg, H, W)
        sd = synth.synth_state_dict(m, seed=wseed)
        m.load_state_dict(sd)
        batch = synth.synth_nuscenes_batch(H, W, b=1, n_cams=n_cams, seed=iseed)
        with torch.no_grad():
            out = m(copy.deepcopy(batch))
        torch.save({"case": name, "kind": "cvt", "config": cfg, "H": H, "W": W, "n_cams": n_cams,
                    "weight_seed": wseed, "input_seed": iseed, "outputs": {k: v.clone() for k, v in out.items()},
                    "weights_checksum": float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point())),
                    "torch": torch.__version__}, os.path.join(HERE, name + ".pt"))
        print(name, {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
