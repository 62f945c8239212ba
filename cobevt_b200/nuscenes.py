"""Drop-in nn.Module surface of the nuScenes SinBEVT fork (nuscenes/cross_view_transformer/model):
PyramidAxialEncoder encoder_pyramid_axial.py:475-558 (Normalize :41-49), Decoder / DecoderBlock
decoder.py:6-61, CrossViewTransformer cvt.py:4-40 -- same constructor arguments, forward contracts
and state_dict keys, forward passes on the sm_100a kernels (inference only, no CPU fallback).

The FAX classes are shared with the OPV2V side (the reference files are byte-identical copies).
``ResNetFeatureBackbone`` is the torchvision-ResNet adaptor of SURVEY 8d (ii): levels 1..3 of a ResNet
(64/128/256 channels at strides 4/8/16) behind the ``backbone=`` injection point, for the boxes where
``efficientnet_pytorch`` (the reference's EfficientNet-B4 prefix) is not installed.
"""
import math
from typing import List

import torch
import torch.nn as nn

from . import engine, ops
from .modules import (_Planned, _require_inference, _ResNetTrunk, BEVEmbedding, Bottleneck,
                      CrossViewSwapAttention)


class Normalize(nn.Module):
    def __init__(self, mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]):
        super().__init__()
        self.register_buffer('mean', torch.tensor(mean)[None, :, None, None], persistent=False)
        self.register_buffer('std', torch.tensor(std)[None, :, None, None], persistent=False)
        self.mean_host, self.std_host = [float(v) for v in mean], [float(v) for v in std]   # no D2H sync per forward


class ResNetFeatureBackbone(_Planned):
    """ResNet levels 1..3 as a CVT backbone: forward(x (N,3,H,W)) -> [ (N,64,H/4,W/4), (N,128,H/8,W/8),
    (N,256,H/16,W/16) ]; ``output_shapes`` as PyramidAxialEncoder expects (:497-504)."""

    def __init__(self, image_height, image_width, num_layers=34):
        super().__init__()
        self.trunk = _ResNetTrunk(num_layers)
        ch = self.trunk.level_channels
        c4 = lambda v, s: (v + s - 1) // s
        self.output_shapes = [torch.Size((1, ch[i], c4(image_height, 4 * 2 ** i), c4(image_width, 4 * 2 ** i)))
                              for i in range(3)]

    def _build_plan(self, dev, dt):
        return engine.ResNetPlan(self.trunk, dev, dt, 3)

    def features_nhwc(self, image_nchw, mean=None, std=None):
        _require_inference(self, image_nchw)
        return self.plan().run(image_nchw.float().contiguous(), nchw=True, mean=mean, std=std)

    def forward(self, x):
        """x is the already normalised image (the encoder applies Normalize before calling the backbone)."""
        return [ops.nhwc_to_nchw(f, torch.float32) for f in self.features_nhwc(x)]


# (kernel, stride, expand_ratio, in, out) of MBConv blocks 0..10 of EfficientNet-B4 (width 1.4, depth 1.8), grouped as
# MODELS['efficientnet-b4'] slices them (backbones/efficientnet.py:15-20)
_EFFNET_B4_BLOCKS = (
    ((3, 1, 1, 48, 24), (3, 1, 1, 24, 24), (3, 2, 6, 24, 32)),
    ((3, 1, 6, 32, 32), (3, 1, 6, 32, 32), (3, 1, 6, 32, 32), (5, 2, 6, 32, 56)),
    ((5, 1, 6, 56, 56), (5, 1, 6, 56, 56), (5, 1, 6, 56, 56), (3, 2, 6, 56, 112)),
    # blocks 11..22 ('reduction_4' proper) are never run by the reference's off-by-one idx_pick (efficientnet.py:46-72)
)
_EFFNET_LAYERS = ['reduction_1', 'reduction_2', 'reduction_3', 'reduction_4']


class _MBConvBlock(nn.Module):
    """Parameter layout of efficientnet_pytorch.model.MBConvBlock (names as in the package / the shipped checkpoint)."""

    def __init__(self, k, stride, expand, cin, cout):
        super().__init__()
        mid = cin * expand
        if expand != 1:
            self._expand_conv = nn.Conv2d(cin, mid, 1, bias=False)
            self._bn0 = nn.BatchNorm2d(mid, momentum=0.01, eps=1e-3)
        self._depthwise_conv = nn.Conv2d(mid, mid, k, stride, groups=mid, bias=False)
        self._bn1 = nn.BatchNorm2d(mid, momentum=0.01, eps=1e-3)
        sq = max(1, int(cin * 0.25))
        self._se_reduce = nn.Conv2d(mid, sq, 1)
        self._se_expand = nn.Conv2d(sq, mid, 1)
        self._project_conv = nn.Conv2d(mid, cout, 1, bias=False)
        self._bn2 = nn.BatchNorm2d(cout, momentum=0.01, eps=1e-3)


class EfficientNetExtractor(_Planned):
    """backbones/efficientnet.py:24-94 on the sm_100a kernels: stem + MBConv blocks 0..10 of EfficientNet-B4, returned per
    ``layer_names`` with the reference's off-by-one pick (``['reduction_2','reduction_3','reduction_4']`` -> the outputs of
    block groups (0,3), (3,7), (7,11): 32 / 56 / 112 channels at strides 4 / 8 / 16).  ``state_dict`` keys equal the
    reference's (``layers.0.{0,1}.*``, ``layers.{1,2,3}.{j}._*``), so the shipped checkpoint loads strictly.
    No weight download: ``EfficientNet.from_pretrained`` becomes random init + ``load_state_dict``.

    The block arithmetic is efficientnet_pytorch==0.7.1's (package absent from the reference tree: parity unpinned,
    see oracle/cobevt_oracle.py): BatchNorm eps 1e-3, swish, SE ratio 0.25 of the block input, and static 'same' padding
    computed for the model's nominal 380-pixel resolution, not for the actual input."""

    nominal_image_size = 380

    def __init__(self, layer_names, image_height, image_width, model_name='efficientnet-b4'):
        super().__init__()
        if model_name != 'efficientnet-b4':
            raise NotImplementedError("only the EfficientNet-B4 prefix the shipped configs use is on the kernel path")
        assert all(k in _EFFNET_LAYERS for k in layer_names)
        idx_max, layer_to_idx = -1, {}
        for i, name in enumerate(_EFFNET_LAYERS):
            if name in layer_names:
                idx_max = max(idx_max, i)
                layer_to_idx[name] = i
        self.block_specs = _EFFNET_B4_BLOCKS[:idx_max]                          # "for idx in range(idx_max)" (:62)
        stem = nn.Sequential(nn.Conv2d(3, 48, 3, 2, bias=False), nn.BatchNorm2d(48, momentum=0.01, eps=1e-3), nn.Identity())
        groups = [nn.Sequential(*[_MBConvBlock(*spec) for spec in grp]) for grp in self.block_specs]
        self.layers = nn.Sequential(stem, *groups)
        self.layer_names = layer_names
        self.idx_pick = [layer_to_idx[l] for l in layer_names]
        # shapes of [stem, group outputs...] for this input size
        nominal = self.nominal_image_size
        def out(sz, k, s, nom):
            b, a = engine.static_same_pad(nom, k, s)
            return (sz + b + a - k) // s + 1
        h, w = out(image_height, 3, 2, nominal), out(image_width, 3, 2, nominal)
        nominal = -(-nominal // 2)
        shapes = [torch.Size((1, 48, h, w))]
        for grp in self.block_specs:
            for (k, s, e, cin, cout) in grp:
                h, w = out(h, k, s, nominal), out(w, k, s, nominal)
                nominal = -(-nominal // s)
            shapes.append(torch.Size((1, grp[-1][4], h, w)))
        if max(self.idx_pick) >= len(shapes):
            raise IndexError("layer_names pick a group the reference never builds")
        self.output_shapes = [shapes[i] for i in self.idx_pick]

    def _build_plan(self, dev, dt):
        return engine.EffNetPlan(self, dev, dt)

    def features_nhwc(self, image_nchw, mean=None, std=None):
        _require_inference(self, image_nchw)
        if 0 in self.idx_pick:
            raise NotImplementedError("picking the stem output ('reduction_1') is not wired on the kernel path")
        feats = self.plan().run(image_nchw.float().contiguous(), mean=mean, std=std)
        return [feats[i - 1] for i in self.idx_pick]

    def forward(self, x):
        return [ops.nhwc_to_nchw(f, torch.float32) for f in self.features_nhwc(x)]


class PyramidAxialEncoder(_Planned):
    def __init__(self, backbone, cross_view: dict, cross_view_swap: dict, bev_embedding: dict, self_attn: dict,
                 dim: list, middle: List[int] = [2, 2], scale: float = 1.0):
        super().__init__()
        self.norm = Normalize()
        self.backbone = backbone
        # scale < 1: F.interpolate(x, scale_factor=scale, recompute_scale_factor=False), i.e. NEAREST down-scaling of
        # every backbone feature map (encoder_pyramid_axial.py:492-495,541)
        self.scale = float(scale)
        assert len(self.backbone.output_shapes) == len(middle)
        cross_views, layers, downs = [], [], []
        for i, (shape, num_layers) in enumerate(zip(self.backbone.output_shapes, middle)):
            _, feat_dim, feat_h, feat_w = shape
            if self.scale < 1.0:
                feat_h, feat_w = int(math.floor(feat_h * self.scale)), int(math.floor(feat_w * self.scale))
            cross_views.append(CrossViewSwapAttention(feat_h, feat_w, feat_dim, dim[i], i, **cross_view, **cross_view_swap))
            layers.append(nn.Sequential(*[Bottleneck(dim[i], dim[i] // 4) for _ in range(num_layers)]))
            if i < len(middle) - 1:
                downs.append(nn.Sequential(nn.Sequential(
                    nn.Conv2d(dim[i], dim[i] // 2, 3, 1, 1, bias=False),
                    nn.PixelUnshuffle(2),
                    nn.Conv2d(dim[i + 1], dim[i + 1], 3, padding=1, bias=False),
                    nn.BatchNorm2d(dim[i + 1]),
                    nn.ReLU(inplace=True),
                    nn.Conv2d(dim[i + 1], dim[i + 1], 1, padding=0, bias=False),
                    nn.BatchNorm2d(dim[i + 1]))))
        self.bev_embedding = BEVEmbedding(dim[0], **bev_embedding)
        self.cross_views = nn.ModuleList(cross_views)
        self.layers = nn.ModuleList(layers)
        self.downsample_layers = nn.ModuleList(downs)

    def _build_plan(self, dev, dt):
        return engine.PyramidPlan(self, dev, dt)

    def run_nhwc(self, batch):
        image = batch['image']
        _require_inference(self, image)
        b, n = image.shape[:2]
        flat = image.flatten(0, 1)
        mean, std = self.norm.mean_host, self.norm.std_host
        if hasattr(self.backbone, 'features_nhwc'):
            feats = self.backbone.features_nhwc(flat, mean=mean, std=std)
        else:   # arbitrary injected torch backbone (e.g. the EfficientNet extractor): its own eager forward
            with torch.no_grad():
                outs = self.backbone((flat - self.norm.mean) / self.norm.std)
            feats = [ops.nchw_to_nhwc(o, self.compute_dtype) for o in outs]
        if self.scale < 1.0:
            feats = [ops.resize_nearest(f, self.scale) for f in feats]
        I = batch['intrinsics'].reshape(b * n, 3, 3).float().contiguous()
        E = batch['extrinsics'].reshape(b * n, 4, 4).float().contiguous()
        return self.plan().run(feats, n, I, E)

    def forward(self, batch):
        return ops.nhwc_to_nchw(self.run_nhwc(batch), torch.float32)


class DecoderBlock(nn.Module):
    def __init__(self, in_channels, out_channels, skip_dim, residual, factor):
        super().__init__()
        dim = out_channels // factor
        self.conv = nn.Sequential(
            nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True),
            nn.Conv2d(in_channels, dim, 3, padding=1, bias=False),
            nn.BatchNorm2d(dim),
            nn.ReLU(inplace=True),
            nn.Conv2d(dim, out_channels, 1, padding=0, bias=False),
            nn.BatchNorm2d(out_channels))
        self.up = nn.Conv2d(skip_dim, out_channels, 1) if residual else None
        self.relu = nn.ReLU(inplace=True)


class Decoder(_Planned):
    def __init__(self, dim, blocks, residual=True, factor=2):
        super().__init__()
        layers, channels = [], dim
        for out_channels in blocks:
            layers.append(DecoderBlock(channels, out_channels, dim, residual, factor))
            channels = out_channels
        self.layers = nn.Sequential(*layers)
        self.out_channels = channels

    def _build_plan(self, dev, dt):
        return engine.CvtDecoderPlan(self, dev, dt)

    def forward(self, x):
        _require_inference(self, x)
        y = self.plan().run(ops.nchw_to_nhwc(x, self.compute_dtype))
        return ops.nhwc_to_nchw(y, torch.float32)


class CrossViewTransformer(_Planned):
    def __init__(self, encoder, decoder, dim_last: int = 64, outputs: dict = {'bev': [0, 1]}):
        super().__init__()
        dim_total = dim_max = 0
        for _, (start, stop) in outputs.items():
            assert start < stop
            dim_total += stop - start
            dim_max = max(dim_max, stop)
        assert dim_max == dim_total
        self.encoder, self.decoder, self.outputs = encoder, decoder, outputs
        self.to_logits = nn.Sequential(
            nn.Conv2d(self.decoder.out_channels, dim_last, 3, padding=1, bias=False),
            nn.BatchNorm2d(dim_last),
            nn.ReLU(inplace=True),
            nn.Conv2d(dim_last, dim_max, 1))

    def _build_plan(self, dev, dt):
        return engine.LogitsPlan(self.to_logits, dev, dt)

    def forward(self, batch):
        x = self.encoder.run_nhwc(batch)
        y = self.decoder.plan().run(x)
        z = self.plan().run(y)
        return {k: z[:, start:stop] for k, (start, stop) in self.outputs.items()}
