"""Import the UNMODIFIED reference from /root/reference (authoring container only).

shapely / matplotlib are not installed and never used by the hot path; empty stub modules
let ``opencood.models.corpbevt`` import."""
import os
import sys
import types

REF_OPV2V = "/root/reference/opv2v"
REF_NUSC = "/root/reference/nuscenes"


def available():
    return os.path.isdir(os.path.join(REF_OPV2V, "opencood"))


def load_opv2v():
    if not available():
        raise RuntimeError("reference not present")
    for name in ("shapely", "shapely.geometry", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["shapely.geometry"].Polygon = object
    if REF_OPV2V not in sys.path:
        sys.path.insert(0, REF_OPV2V)
    import opencood.models.corpbevt as corp
    import opencood.models.fax_fused_transformer as fft
    import opencood.models.sub_modules.fax_modules as fax
    import opencood.models.fusion_modules.swap_fusion_modules as swap
    import opencood.models.backbones.resnet_ms as resnet_ms
    return types.SimpleNamespace(corpbevt=corp, fax_fused=fft, fax=fax, swap=swap, resnet_ms=resnet_ms)


def load_nuscenes():
    """nuscenes/cross_view_transformer model modules (pure torch + einops; no hydra/lightning needed)."""
    if not os.path.isdir(os.path.join(REF_NUSC, "cross_view_transformer")):
        raise RuntimeError("reference not present")
    if REF_NUSC not in sys.path:
        sys.path.insert(0, REF_NUSC)
    import cross_view_transformer.model.encoder_pyramid_axial as epa
    import cross_view_transformer.model.decoder as dec
    import cross_view_transformer.model.cvt as cvt
    return types.SimpleNamespace(epa=epa, decoder=dec, cvt=cvt)
