"""Drop-in nn.Module surface of the nuScenes SinBEVT fork (nuscenes/cross_view_transformer/model):
PyramidAxialEncoder encoder_pyramid_axial.py:475-558 (Normalize :41-49), Decoder / DecoderBlock
decoder.py:6-61, CrossViewTransformer cvt.py:4-40 -- same constructor arguments, forward contracts
and state_dict keys, forward passes on the sm_100a kernels (inference only, no CPU fallback).

The FAX classes are shared with the OPV2V side (the reference files are byte-identical copies).
``ResNetFeatureBackbone`` is the torchvision-ResNet adaptor of SURVEY 8d (ii): levels 1..3 of a ResNet
(64/128/256 channels at strides 4/8/16) behind the ``backbone=`` injection point, for the boxes where
``efficientnet_pytorch`` (the reference's EfficientNet-B4 prefix) is not installed.
"""
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


class PyramidAxialEncoder(_Planned):
    def __init__(self, backbone, cross_view: dict, cross_view_swap: dict, bev_embedding: dict, self_attn: dict,
                 dim: list, middle: List[int] = [2, 2], scale: float = 1.0):
        super().__init__()
        self.norm = Normalize()
        self.backbone = backbone
        if scale < 1.0:
            raise NotImplementedError("scale < 1 (feature down-scaling) is not on the accelerated path")
        assert len(self.backbone.output_shapes) == len(middle)
        cross_views, layers, downs = [], [], []
        for i, (shape, num_layers) in enumerate(zip(self.backbone.output_shapes, middle)):
            _, feat_dim, feat_h, feat_w = shape
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
        mean = self.norm.mean.flatten().tolist()
        std = self.norm.std.flatten().tolist()
        if hasattr(self.backbone, 'features_nhwc'):
            feats = self.backbone.features_nhwc(flat, mean=mean, std=std)
        else:   # arbitrary injected torch backbone (e.g. the EfficientNet extractor): its own eager forward
            with torch.no_grad():
                outs = self.backbone((flat - self.norm.mean) / self.norm.std)
            feats = [ops.nchw_to_nhwc(o, self.compute_dtype) for o in outs]
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
