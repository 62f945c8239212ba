"""Generates tests/golden/*.pt from the UNMODIFIED reference (run in the authoring container,
where /root/reference exists):

    python tests/golden/make_golden.py                 # every fixture that is missing
    python tests/golden/make_golden.py --force NAME    # regenerate one
    python tests/golden/make_golden.py --autocast      # tests/golden/autocast_error.json (see below)

Each fixture holds only the seeded-input recipe and the REFERENCE's output; weights and inputs
are regenerated bit-identically from cobevt_b200.synth on any box, so the GPU box (which has no
/root/reference) can compare the CUDA path and the oracle against what the reference produced.
The one exception is `cvt_realweights_6cam`: the trained (non-backbone) weights of the checkpoint
shipped with the reference (nuscenes/logs/sinbevt_nuscenes_vehicle_50k.ckpt), matrices rounded to
fp16, are stored in the fixture, and the reference output was produced WITH those rounded weights.

`--autocast` measures how far the reference itself moves when it is run in 16-bit: for every
fixture the unmodified reference forward under torch.autocast('cpu', float16 / bfloat16) is
compared with its own fp32 forward (same metric as the tests: max|d| / max|ref|).  These numbers
are the yard-stick for the whole-model tolerances in tests/test_golden.py / test_model_gpu.py.
"""
import argparse
import copy
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refimport  # noqa: E402
from cobevt_b200 import configs, synth  # noqa: E402


def _nomask(cfg):
    cfg["fax_fusion"]["mask"] = False
    return cfg


def _compressed(cfg, rate):
    cfg["compression"] = rate
    return cfg


CASES = {
    # name: (config builder, model kind, record_len, n_cams, weight seed, input seed)
    "corpbevt_small_ragged": (configs.small, "corpbevt", (3, 2), 2, 0, 1),
    "corpbevt_small_5x4": (configs.small, "corpbevt", (5,), 4, 0, 1),
    "corpbevt_small_nomask": (lambda: _nomask(configs.small()), "corpbevt", (2, 5), 2, 0, 3),
    "corpbevt_small_compress": (lambda: _compressed(configs.small(), 4), "corpbevt", (2, 3), 2, 0, 7),
    "faxfused_c1_224": (lambda: configs.fax_fused_args(configs.corpbevt_args(
        image_h=224, image_w=224, bev=224, q_win=((16, 16), (8, 8), (28, 28)), feat_win=((4, 4), (2, 2), (7, 7)),
        self_attn_window=28)), "fax", (1,), 1, 0, 4),
    "corpbevt_northstar_1agent": (configs.north_star, "corpbevt", (1,), 6, 0, 5),
    "corpbevt_northstar_2agents": (configs.north_star, "corpbevt", (2,), 6, 0, 6),
    # BASELINE C3/C4 geometry exactly as corpbevt.yaml ships it: 4 cams x 512x512, ResNet34 id_pick [1,2,3]
    # (layer4 runs), 8x8 / 8x8 / 16x16 feature windows
    "corpbevt_native_1agent_4cam": (configs.native_opv2v, "corpbevt", (1,), 4, 0, 8),
    # the metric's full sample (5 agents x 6 cams, non-identity agent transforms) inside a ragged batch
    "corpbevt_northstar_ragged_5_3_1_4": (configs.north_star, "corpbevt", (5, 3, 1, 4), 6, 0, 9),
}


NUSC_CASES = {
    # name: (config builder, H, W, n_cams, weight seed, input seed)
    "cvt_native200_1cam": (lambda: configs.nuscenes_cvt(224, 224, 200, feat_win=((6, 6), (6, 6), (14, 14))), 224, 224, 1, 0, 2),
    "cvt_c2_bev256_2cam": (configs.nuscenes_c2, 224, 480, 2, 0, 2),
    "cvt_c2_bev256_6cam": (configs.nuscenes_c2, 224, 480, 6, 0, 4),        # BASELINE C2: all six cameras
}


def _opv2v_case(ref, name):
    mk, kind, record_len, n_cams, wseed, iseed = CASES[name]
    cfg = mk()
    cls = ref.corpbevt.CorpBEVT if kind == "corpbevt" else ref.fax_fused.FaxFusedTransformer
    m = cls(copy.deepcopy(cfg)).eval()
    sd = synth.synth_state_dict(m, seed=wseed)
    m.load_state_dict(sd)
    batch = synth.synth_batch(dict(cfg, max_cav=cfg.get("max_cav", 1)), record_len=record_len, n_cams=n_cams, seed=iseed)
    if kind == "fax":
        batch["inputs"] = batch["inputs"].reshape(len(record_len), 1, *batch["inputs"].shape[2:])
    return cfg, kind, m, sd, batch


def _nusc_case(nref, name):
    import nusc_common as nc
    mk, H, W, n_cams, wseed, iseed = NUSC_CASES[name]
    cfg = mk()
    m = nc.build_reference(nref, cfg, H, W)
    sd = synth.synth_state_dict(m, seed=wseed)
    m.load_state_dict(sd)
    batch = synth.synth_nuscenes_batch(H, W, b=1, n_cams=n_cams, seed=iseed)
    return cfg, m, sd, batch


def _checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))


REAL = "cvt_realweights_6cam"


def _real_case(nref):
    """Trained weights of the shipped checkpoint (everything but the EfficientNet backbone, whose package is
    absent) on the reference-native 200x200 config, 6 cameras; a deterministic stand-in produces the three
    backbone feature maps (32/56/112 channels), regenerated identically on the test side."""
    import nusc_common as nc
    H, W = 224, 480
    cfg = configs.nuscenes_cvt()
    e = copy.deepcopy(cfg['encoder'])
    enc = nref.epa.PyramidAxialEncoder(nc.FakeEffNetFeatures(H, W), e['cross_view'], e['cross_view_swap'],
                                       e['bev_embedding'], e['self_attn'], e['dim'], e['middle'], e['scale'])
    m = nref.cvt.CrossViewTransformer(enc, nref.decoder.Decoder(128, [128, 128, 64], True, 2), 64,
                                      {'bev': [0, 1], 'center': [1, 2]}).eval()
    ck = {k[len("backbone."):]: v for k, v in nc.load_shipped_ckpt().items() if k.startswith("backbone.")}
    ck = {k: v for k, v in ck.items() if not k.startswith("encoder.backbone.")}
    ck = {k: (v.half().float() if (v.is_floating_point() and v.dim() >= 2) else v) for k, v in ck.items()}
    missing, unexpected = m.load_state_dict(ck, strict=False)
    assert not unexpected and all(k.startswith("encoder.backbone") for k in missing)
    batch = synth.synth_nuscenes_batch(H, W, b=1, n_cams=6, seed=3)
    stored = {k: (v.half() if (v.is_floating_point() and v.dim() >= 2) else v) for k, v in ck.items()}
    return cfg, m, stored, batch, H, W


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", nargs="*", default=[])
    ap.add_argument("--autocast", action="store_true")
    args = ap.parse_args()
    ref = refimport.load_opv2v()
    nref = refimport.load_nuscenes()
    ac = {}
    ac_path = os.path.join(HERE, "autocast_error.json")
    if args.autocast and os.path.exists(ac_path):
        ac = json.load(open(ac_path))

    def autocast_err(name, m, batch, pick):
        if name in ac and name not in args.force:
            return
        with torch.no_grad():
            want = pick(m(copy.deepcopy(batch)))
            row = {}
            for tag, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
                with torch.autocast("cpu", dtype=dt):
                    got = pick(m(copy.deepcopy(batch)))
                row[tag] = {k: _rel(got[k], want[k]) for k in want}
        ac[name] = row
        print("autocast", name, row, flush=True)
        json.dump(ac, open(ac_path, "w"), indent=1, sort_keys=True)

    for name in CASES:
        path = os.path.join(HERE, name + ".pt")
        need = not os.path.exists(path) or name in args.force
        if not need and not args.autocast:
            continue
        cfg, kind, m, sd, batch = _opv2v_case(ref, name)
        if need:
            with torch.no_grad():
                out = m(copy.deepcopy(batch))
            dyn = out["dynamic_seg"]
            torch.save({"case": name, "kind": kind, "config": cfg, "record_len": CASES[name][2], "n_cams": CASES[name][3],
                        "weight_seed": CASES[name][4], "input_seed": CASES[name][5], "dynamic_seg": dyn.clone(),
                        "weights_checksum": _checksum(sd), "torch": torch.__version__}, path)
            print(name, tuple(dyn.shape), float(dyn.abs().max()), flush=True)
        if args.autocast:
            autocast_err(name, m, batch, lambda o: {"dynamic_seg": o["dynamic_seg"]})
    for name in NUSC_CASES:
        path = os.path.join(HERE, name + ".pt")
        need = not os.path.exists(path) or name in args.force
        if not need and not args.autocast:
            continue
        cfg, m, sd, batch = _nusc_case(nref, name)
        if need:
            with torch.no_grad():
                out = m(copy.deepcopy(batch))
            torch.save({"case": name, "kind": "cvt", "config": cfg, "H": NUSC_CASES[name][1], "W": NUSC_CASES[name][2],
                        "n_cams": NUSC_CASES[name][3], "weight_seed": NUSC_CASES[name][4], "input_seed": NUSC_CASES[name][5],
                        "outputs": {k: v.clone() for k, v in out.items()},
                        "weights_checksum": _checksum(sd), "torch": torch.__version__}, path)
            print(name, {k: tuple(v.shape) for k, v in out.items()}, flush=True)
        if args.autocast:
            autocast_err(name, m, batch, lambda o: dict(o))
    path = os.path.join(HERE, REAL + ".pt")
    need = not os.path.exists(path) or REAL in args.force
    if need or args.autocast:
        cfg, m, stored, batch, H, W = _real_case(nref)
        if need:
            with torch.no_grad():
                out = m(copy.deepcopy(batch))
            torch.save({"case": REAL, "kind": "cvt_real", "config": cfg, "H": H, "W": W, "n_cams": 6, "input_seed": 3,
                        "weights": stored, "outputs": {k: v.clone() for k, v in out.items()},
                        "torch": torch.__version__}, path)
            print(REAL, {k: (tuple(v.shape), float(v.abs().max())) for k, v in out.items()}, flush=True)
        if args.autocast:
            autocast_err(REAL, m, batch, lambda o: dict(o))


if __name__ == "__main__":
    main()
