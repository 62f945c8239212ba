"""Shared helpers for the nuScenes (CVT fork) tests: build the reference / our CrossViewTransformer with the
torchvision-ResNet34 backbone adaptor from one config dict."""
import copy

import torch
import torch.nn as nn


class TorchResNetBackbone(nn.Module):
    """Reference-side adaptor: torchvision resnet34 levels 1..3 behind PyramidAxialEncoder's backbone= slot.
    Parameter names mirror cobevt_b200.nuscenes.ResNetFeatureBackbone ('trunk.*')."""

    def __init__(self, image_height, image_width):
        super().__init__()
        import torchvision
        self.trunk = torchvision.models.resnet34(weights=None)
        c4 = lambda v, s: (v + s - 1) // s
        self.output_shapes = [torch.Size((1, c, c4(image_height, s), c4(image_width, s)))
                              for c, s in ((64, 4), (128, 8), (256, 16))]

    def forward(self, x):
        t = self.trunk
        x = t.maxpool(t.relu(t.bn1(t.conv1(x))))
        x1 = t.layer1(x)
        x2 = t.layer2(x1)
        x3 = t.layer3(x2)
        return [x1, x2, x3]


def build_reference(ref, cfg, H, W):
    e = copy.deepcopy(cfg['encoder'])
    enc = ref.epa.PyramidAxialEncoder(TorchResNetBackbone(H, W), e['cross_view'], e['cross_view_swap'],
                                      e['bev_embedding'], e['self_attn'], e['dim'], e['middle'], e['scale'])
    dec = ref.decoder.Decoder(cfg['decoder_dim'], cfg['decoder_blocks'], True, 2)
    return ref.cvt.CrossViewTransformer(enc, dec, cfg['dim_last'], copy.deepcopy(cfg['outputs'])).eval()


def build_ours(cfg, H, W):
    from cobevt_b200 import nuscenes as ns
    e = copy.deepcopy(cfg['encoder'])
    enc = ns.PyramidAxialEncoder(ns.ResNetFeatureBackbone(H, W), e['cross_view'], e['cross_view_swap'],
                                 e['bev_embedding'], e['self_attn'], e['dim'], e['middle'], e['scale'])
    dec = ns.Decoder(cfg['decoder_dim'], cfg['decoder_blocks'], True, 2)
    return ns.CrossViewTransformer(enc, dec, cfg['dim_last'], copy.deepcopy(cfg['outputs'])).eval()


def oracle_forward(batch, sd, cfg):
    from oracle import cobevt_oracle as orc
    backbone = lambda img: orc.resnet_levels(img, sd, "encoder.backbone.trunk.", 34, 3)
    return orc.cross_view_transformer(batch, sd, cfg, backbone)


class FakeEffNetFeatures(nn.Module):
    """Shape-compatible, deterministic stand-in for the EfficientNet-B4 prefix (32/56/112 channels at strides
    4/8/16) behind the ``backbone=`` injection point: real-weight parity of everything after the backbone on
    boxes where ``efficientnet_pytorch`` is absent."""

    def __init__(self, H, W):
        super().__init__()
        self.output_shapes = [torch.Size((1, c, H // s, W // s)) for c, s in ((32, 4), (56, 8), (112, 16))]

    def forward(self, x):
        g = torch.Generator().manual_seed(7)
        return [torch.randn(x.shape[0], *s[1:], generator=g).to(x.device) for s in self.output_shapes]


def load_shipped_ckpt(path="/root/reference/nuscenes/logs/sinbevt_nuscenes_vehicle_50k.ckpt"):
    """state_dict of the Lightning checkpoint shipped with the reference: omegaconf /
    lightning classes inside the pickle are replaced by permissive stubs."""
    import pickle
    import types

    class _Stub:
        def __init__(self, *a, **k):
            pass

        def __setstate__(self, st):
            self.__dict__.update(st if isinstance(st, dict) else {})

    class _U(pickle.Unpickler):
        def find_class(self, mod, name):
            try:
                return super().find_class(mod, name)
            except Exception:
                return type(name, (_Stub,), {})

    pm = types.ModuleType("stub_pickle")
    pm.Unpickler = _U
    pm.load = lambda f, **kw: _U(f).load()
    pm.__name__ = "pickle"
    return torch.load(path, map_location="cpu", weights_only=False, pickle_module=pm)["state_dict"]
