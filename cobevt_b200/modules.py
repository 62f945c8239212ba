"""Drop-in nn.Module surface of the CoBEVT FAX path (OPV2V side).

Every class keeps the reference's constructor arguments, forward contract and
``state_dict`` keys/shapes so reference checkpoints load unchanged; the forward passes run
the sm_100a kernels through ``engine``/``ops``.  Parameters are ordinary fp32
``nn.Parameter``s; folded/packed 16-bit weight images are derived caches ("plans") rebuilt
after ``load_state_dict`` / device moves and never saved.

Inference only: a forward in training mode or on a non-CUDA tensor raises -- there is no CPU or
eager fallback.  The kernels record no autograd graph: outputs are plain (detached) tensors
whether or not grad mode is on, exactly like a forward under ``torch.no_grad()``.

Shape violations raise what the reference raises: ``einops.EinopsError`` where the reference's
``rearrange`` would reject a non-divisible window, ``AssertionError`` where it asserts.

Reference classes mirrored (paths under opv2v/opencood/models):
  backbones/resnet_ms.py ResnetEncoder :8-89; sub_modules/fax_modules.py BEVEmbedding :38-90,
  Attention :93-176, CrossWinAttention :179-248, CrossViewSwapAttention :251-441,
  FAXModule :444-521; fusion_modules/swap_fusion_modules.py Attention :13-128,
  SwapFusionBlockMask :131-192, SwapFusionBlock :195-230, SwapFusionEncoder :233-286;
  base_transformer.py PreNormResidual :102-109, FeedForward :112-124;
  sub_modules/naive_decoder.py :8-91; sub_modules/bev_seg_head.py :10-63;
  corpbevt.py STTF :22-64, CorpBEVT :67-145; fax_fused_transformer.py :13-48.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from . import engine, ops

COMPUTE_DTYPE = torch.float16


def _einops_error(msg):
    """The exception type the reference's einops.rearrange raises for a non-divisible partition."""
    try:
        from einops import EinopsError
    except Exception:                      # einops absent: its error derives from RuntimeError
        return RuntimeError(msg)
    return EinopsError(msg)


def set_compute_dtype(module, dtype):
    """Select the 16-bit storage / tensor-core operand type (torch.float16 or torch.bfloat16; fp32
    accumulation either way) for every kernel-backed sub-module of ``module``; plans are rebuilt lazily."""
    if dtype not in (torch.float16, torch.bfloat16):
        raise TypeError(f"compute dtype must be float16 or bfloat16, got {dtype}")
    for m in module.modules():
        if isinstance(m, _Planned):
            m.compute_dtype = dtype
            object.__setattr__(m, "_plan", None)
    return module


def _require_inference(mod, *tensors):
    if mod.training:
        raise RuntimeError(f"{type(mod).__name__}: the B200 kernels are forward/inference only; call .eval()")
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError(f"{type(mod).__name__}: input is on {t.device}; there is no CPU path")


class _Planned(nn.Module):
    """Mixin: lazily built, automatically invalidated kernel plan."""

    compute_dtype = COMPUTE_DTYPE

    def __init__(self):
        super().__init__()
        object.__setattr__(self, "_plan", None)
        self.register_load_state_dict_post_hook(lambda m, _: m.invalidate_plan())

    def invalidate_plan(self):
        for m in self.modules():
            if isinstance(m, _Planned):
                object.__setattr__(m, "_plan", None)

    def _apply(self, fn, *a, **k):
        self.invalidate_plan()
        return super()._apply(fn, *a, **k)

    def plan(self):
        if self._plan is None:
            dev = next(self.parameters()).device
            with torch.no_grad():
                object.__setattr__(self, "_plan", self._build_plan(dev, self.compute_dtype))
        return self._plan


# --------------------------------------------------------------------------- ResNet trunk
class _BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, cin, planes, stride, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample


class Bottleneck(_Planned):
    """torchvision.models.resnet.Bottleneck (v1.5: stride on conv2) as the FAX stages use it
    (fax_modules.py:10,472: ``Bottleneck(c, c // 4)``, stride 1, identity shortcut)."""
    expansion = 4

    def __init__(self, cin, planes, stride=1, downsample=None):
        super().__init__()
        self.stride = stride
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def _build_plan(self, dev, dt):
        return engine.BottleneckPlan(self, dev, dt)

    def forward(self, x):
        """x (N, C, H, W) -> (N, C, H, W)."""
        _require_inference(self, x)
        if self.stride != 1 or self.downsample is not None:
            raise NotImplementedError("only the stride-1 identity-shortcut Bottleneck of the FAX stages is on the kernel path")
        y = self.plan().run(ops.nchw_to_nhwc(x, self.compute_dtype))
        return ops.nhwc_to_nchw(y, torch.float32)


_RESNET_CFG = {18: (_BasicBlock, (2, 2, 2, 2)), 34: (_BasicBlock, (3, 4, 6, 3)),
               50: (Bottleneck, (3, 4, 6, 3)), 101: (Bottleneck, (3, 4, 23, 3)), 152: (Bottleneck, (3, 8, 36, 3))}


class _ResNetTrunk(nn.Module):
    """Same parameter names as torchvision's ResNet (conv1, bn1, layer1..4, fc)."""

    def __init__(self, num_layers):
        super().__init__()
        block, counts = _RESNET_CFG[num_layers]
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        cin = 64
        for li, (planes, n) in enumerate(zip((64, 128, 256, 512), counts)):
            blocks = []
            for bi in range(n):
                stride = 2 if (li > 0 and bi == 0) else 1
                ds = None
                if stride != 1 or cin != planes * block.expansion:
                    ds = nn.Sequential(nn.Conv2d(cin, planes * block.expansion, 1, stride, bias=False),
                                       nn.BatchNorm2d(planes * block.expansion))
                blocks.append(block(cin, planes, stride, ds))
                cin = planes * block.expansion
            setattr(self, f"layer{li + 1}", nn.Sequential(*blocks))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(cin, 1000)
        self.level_channels = [p * block.expansion for p in (64, 128, 256, 512)]


class ResnetEncoder(_Planned):
    """resnet_ms.py:8-89.  forward(input_images (B,L,M,H,W,3)) -> list/one of (B,L,M,C,h,w)."""

    def __init__(self, params):
        super().__init__()
        self.num_layers = params['num_layers']
        self.pretrained = params['pretrained']
        self.idx_pick = params['id_pick']
        if self.num_layers not in _RESNET_CFG:
            raise ValueError("{} is not a valid number of resnet layers".format(self.num_layers))
        if self.num_layers not in (18, 34):
            raise ValueError(f"ResNet-{self.num_layers} needs Cout up to 2048; the sm_100a conv kernel serves Cout <= 512 "
                             "(BasicBlock ResNet-18/34, the depths the shipped configs use)")
        if self.pretrained:
            raise RuntimeError("pretrained=True needs a weight download; load a checkpoint with load_state_dict")
        self.encoder = _ResNetTrunk(self.num_layers)
        # uint8 camera frames: when forward() receives a uint8 tensor the RgbPreProcessor arithmetic
        # (rgb_preprocessor.py:16-36: optional BGR->RGB, /255, (x - mean) / std) runs inside the stem's staging kernel
        self.input_mean, self.input_std, self.input_bgr = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225), False
        h, w = params['image_height'], params['image_width']
        shapes = []
        for li, c in enumerate(self.encoder.level_channels):
            s = 4 * 2 ** li
            shapes.append(torch.Size((1, 1, 1, c, (h + s - 1) // s, (w + s - 1) // s)))
        self.output_shapes = [shapes[i] for i in self.idx_pick] if isinstance(self.idx_pick, list) else shapes[self.idx_pick]

    def _levels(self):
        return (max(self.idx_pick) if isinstance(self.idx_pick, list) else self.idx_pick) + 1

    def _build_plan(self, dev, dt):
        return engine.ResNetPlan(self.encoder, dev, dt, self._levels())

    def features_nhwc(self, input_images):
        """engine-level entry: list (per level, up to the deepest picked) of (B*L*M, h, w, C)."""
        _require_inference(self, input_images)
        b, l, m, h, w, c = input_images.shape
        if input_images.dtype == torch.uint8:
            return self.plan().run(input_images.reshape(b * l * m, h, w, c).contiguous(), mean=self.input_mean,
                                   std=self.input_std, bgr=self.input_bgr)
        return self.plan().run(input_images.reshape(b * l * m, h, w, c).float().contiguous())

    def forward(self, input_images):
        b, l, m = input_images.shape[:3]
        feats = self.features_nhwc(input_images)
        if max(self.idx_pick if isinstance(self.idx_pick, list) else [self.idx_pick]) >= len(feats):
            raise IndexError("id_pick beyond the encoder levels")
        res = [ops.nhwc_to_nchw(f, torch.float32).view(b, l, m, f.shape[3], f.shape[1], f.shape[2]) for f in feats]
        if isinstance(self.idx_pick, list):
            return [res[i] for i in self.idx_pick]
        return res[self.idx_pick]


# --------------------------------------------------------------------------- FAX (SinBEVT)
def _unit_grid(height, width):
    xs = torch.linspace(0, 1, width)
    ys = torch.linspace(0, 1, height)
    return torch.stack([xs[None, :].expand(height, width), ys[:, None].expand(height, width),
                        torch.ones(height, width)], 0)


class BEVEmbedding(nn.Module):
    def __init__(self, dim, sigma, bev_height, bev_width, h_meters, w_meters, offset, upsample_scales):
        super().__init__()
        sh, sw = bev_height / h_meters, bev_width / w_meters
        V_inv = torch.tensor([[0., -sw, bev_width / 2.], [-sh, 0., bev_height * offset + bev_height / 2.],
                              [0., 0., 1.]]).inverse()
        for i, scale in enumerate(upsample_scales):
            h, w = bev_height // scale, bev_width // scale
            g = _unit_grid(h, w)
            g[0] *= bev_width
            g[1] *= bev_height
            self.register_buffer('grid%d' % i, (V_inv @ g.reshape(3, -1)).reshape(3, h, w), persistent=False)
        self.learned_features = nn.Parameter(
            sigma * torch.randn(dim, bev_height // upsample_scales[0], bev_width // upsample_scales[0]))

    def get_prior(self):
        return self.learned_features


class Attention(_Planned):
    """fax_modules.Attention: global rel-pos self-attention over a (window_size x window_size) map."""

    def __init__(self, dim, dim_head=32, dropout=0., window_size=25):
        super().__init__()
        assert (dim % dim_head) == 0, 'dimension should be divisible by dimension per head'
        assert dim_head == 32, "kernels are built for head_dim 32"
        self.heads = dim // dim_head
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, dim * 3, bias=False)
        self.attend = nn.Sequential(nn.Softmax(dim=-1), nn.Dropout(dropout))
        self.to_out = nn.Sequential(nn.Linear(dim, dim, bias=False), nn.Dropout(dropout))
        self.rel_pos_bias = nn.Embedding((2 * window_size - 1) ** 2, self.heads)
        pos = torch.arange(window_size)
        yy, xx = torch.meshgrid(pos, pos, indexing='ij')
        g = torch.stack([yy.reshape(-1), xx.reshape(-1)], 1)
        rel = g[:, None] - g[None, :] + window_size - 1
        self.register_buffer('rel_pos_indices', rel[..., 0] * (2 * window_size - 1) + rel[..., 1], persistent=False)

    def _build_plan(self, dev, dt):
        return engine.GlobalAttnPlan(self, dev, dt)

    def forward(self, x):
        _require_inference(self, x)
        y = self.plan().run(ops.nchw_to_nhwc(x, self.compute_dtype))
        return ops.nhwc_to_nchw(y, x.dtype if x.dtype == torch.float32 else torch.float32)


class CrossWinAttention(_Planned):
    def __init__(self, dim, heads, dim_head, qkv_bias, rel_pos_emb=False, norm=nn.LayerNorm):
        super().__init__()
        assert dim_head == 32, "kernels are built for head_dim 32"
        self.scale = dim_head ** -0.5
        self.heads, self.dim_head, self.rel_pos_emb = heads, dim_head, rel_pos_emb
        self.to_q = nn.Sequential(norm(dim), nn.Linear(dim, heads * dim_head, bias=qkv_bias))
        self.to_k = nn.Sequential(norm(dim), nn.Linear(dim, heads * dim_head, bias=qkv_bias))
        self.to_v = nn.Sequential(norm(dim), nn.Linear(dim, heads * dim_head, bias=qkv_bias))
        self.proj = nn.Linear(heads * dim_head, dim)

    def _build_plan(self, dev, dt):
        return engine.CrossAttnPlan(self, dev, dt)

    def forward(self, q, k, v, skip=None):
        """q (b n X Y W1 W2 d), k/v (b n x y w1 w2 d) -> (b X Y W1 W2 d): the tensors arrive already
        window-partitioned, so they are viewed as stacks of single-window maps."""
        _require_inference(self, q, k, v)
        assert k.shape == v.shape
        b, n, X, Y, W1, W2, d = q.shape
        _, nk, x_, y_, w1, w2, _ = k.shape
        assert X * Y == x_ * y_           # fax_modules.py:208 (q_height * q_width == kv_height * kv_width)
        pl, dt = self.plan(), self.compute_dtype
        tok = lambda t: t.permute(0, 2, 3, 1, 4, 5, 6).reshape(-1, t.shape[1], t.shape[4], t.shape[5], d).to(dt).contiguous()
        qt, kt, vt = tok(q), tok(k), tok(v)                                  # (b*X*Y, n, W1, W2, d)
        qh = ops.linear_tokens(ops.layernorm_rows(qt.reshape(-1, d)), pl.q)[0].view(qt.shape)
        kh = ops.linear_tokens(ops.layernorm_rows(kt.reshape(-1, d)), pl.k)[0].view(kt.shape)
        vh = ops.linear_tokens(ops.layernorm_rows(vt.reshape(-1, d)), pl.v)[0].view(vt.shape)
        a = ops.window_attention(qh, kh, vh, ops.token_map(n, W1, W2, W1, W2, False),
                                 ops.token_map(nk, w1, w2, w1, w2, False), self.heads, self.scale, True)
        res = None
        if skip is not None:
            res = skip.reshape(-1, d).to(dt).contiguous()
        z = ops.linear_tokens(a.reshape(-1, d), pl.proj, res=res, want16=False, out32=True)[1]
        return z.view(b, X, Y, W1, W2, d).to(q.dtype)


class CrossViewSwapAttention(_Planned):
    def __init__(self, feat_height, feat_width, feat_dim, dim, index, image_height, image_width, qkv_bias,
                 q_win_size, feat_win_size, heads, dim_head, bev_embedding_flag, rel_pos_emb=False,
                 no_image_features=False, skip=True, norm=nn.LayerNorm):
        super().__init__()
        plane = _unit_grid(feat_height, feat_width)[None, None]
        plane[:, :, 0] *= image_width
        plane[:, :, 1] *= image_height
        self.register_buffer('image_plane', plane, persistent=False)
        self.image_height, self.image_width = image_height, image_width
        self.feature_linear = nn.Sequential(nn.BatchNorm2d(feat_dim), nn.ReLU(), nn.Conv2d(feat_dim, dim, 1, bias=False))
        self.feature_proj = None if no_image_features else nn.Sequential(
            nn.BatchNorm2d(feat_dim), nn.ReLU(), nn.Conv2d(feat_dim, dim, 1, bias=False))
        self.bev_embed_flag = bev_embedding_flag[index]
        if self.bev_embed_flag:
            self.bev_embed = nn.Conv2d(2, dim, 1)
        self.img_embed = nn.Conv2d(4, dim, 1, bias=False)
        self.cam_embed = nn.Conv2d(4, dim, 1, bias=False)
        self.q_win_size, self.feat_win_size = q_win_size[index], feat_win_size[index]
        self.rel_pos_emb = rel_pos_emb
        self.cross_win_attend_1 = CrossWinAttention(dim, heads[index], dim_head[index], qkv_bias)
        self.cross_win_attend_2 = CrossWinAttention(dim, heads[index], dim_head[index], qkv_bias)
        self.skip = skip
        self.prenorm_1, self.prenorm_2 = norm(dim), norm(dim)
        self.mlp_1 = nn.Sequential(nn.Linear(dim, 2 * dim), nn.GELU(), nn.Linear(2 * dim, dim))
        self.mlp_2 = nn.Sequential(nn.Linear(dim, 2 * dim), nn.GELU(), nn.Linear(2 * dim, dim))
        self.postnorm = norm(dim)
        object.__setattr__(self, "_world", None)

    def _build_plan(self, dev, dt):
        if self.bev_embed_flag and self._world is None:
            raise RuntimeError("stage uses the BEV embedding: call forward(index, x, bev, ...) first")
        return engine.FaxStagePlan(self, self._world, dev, dt)

    def forward(self, index, x, bev, feature, I_inv, E_inv):
        """x (b,d,H,W), feature (b,n,c,h,w), I_inv (b,n,3,3), E_inv (b,n,4,4) -> (b,d,H,W)."""
        _require_inference(self, x, feature)
        if self._world is None:
            object.__setattr__(self, "_world", getattr(bev, 'grid%d' % index)[:2])
            object.__setattr__(self, "_plan", None)
        b, n, c, h, w = feature.shape
        dt = self.compute_dtype
        if x.shape[2] % self.q_win_size[0] or x.shape[3] % self.q_win_size[1]:
            # the reference's rearrange 'b d (x w1) (y w2) -> ...' (fax_modules.py:399-401) rejects this
            raise _einops_error(f"query map {tuple(x.shape[2:])} is not divisible by the window {tuple(self.q_win_size)}")
        xs = ops.nchw_to_nhwc(x, dt)
        f = ops.nchw_to_nhwc(feature.reshape(b * n, c, h, w), dt)
        y = self.plan().run(xs, None, b, f, n, I_inv.reshape(b * n, 3, 3).float().contiguous(),
                            E_inv.reshape(b * n, 4, 4).float().contiguous(), invert_intrinsic=False)
        return ops.nhwc_to_nchw(y, torch.float32)


class FAXModule(_Planned):
    def __init__(self, config):
        super().__init__()
        middle, dim = config['middle'], config['dim']
        self.backbone_output_shape = config['backbone_output_shape']
        assert len(middle) == len(self.backbone_output_shape)
        cross_views, layers, downs = [], [], []
        for i, (shape, num_layers) in enumerate(zip(self.backbone_output_shape, middle)):
            _, _, _, feat_dim, feat_h, feat_w = shape
            cross_views.append(CrossViewSwapAttention(feat_h, feat_w, feat_dim, dim[i], i,
                                                      **config['cross_view'], **config['cross_view_swap']))
            layers.append(nn.Sequential(*[Bottleneck(dim[i], dim[i] // 4) for _ in range(num_layers)]))
            if i < len(middle) - 1:
                downs.append(nn.Sequential(nn.Sequential(
                    nn.Conv2d(dim[i], dim[i] // 4, 3, 1, 1, bias=False),
                    nn.PixelUnshuffle(2),
                    nn.Conv2d(dim[i + 1], dim[i + 1], 3, padding=1, bias=False),
                    nn.BatchNorm2d(dim[i + 1]),
                    nn.ReLU(inplace=True),
                    nn.Conv2d(dim[i + 1], dim[i + 1], 1, padding=0, bias=False),
                    nn.BatchNorm2d(dim[i + 1]))))
        self.bev_embedding = BEVEmbedding(dim[0], **config['bev_embedding'])
        self.cross_views = nn.ModuleList(cross_views)
        self.layers = nn.ModuleList(layers)
        self.downsample_layers = nn.ModuleList(downs)
        self.self_attn = Attention(dim[-1], **config['self_attn'])

    def _build_plan(self, dev, dt):
        return engine.FaxPlan(self, dev, dt)

    def run_nhwc(self, feats, n, intrinsic, extrinsic):
        """engine-level entry: feats list of (BL*n, h, w, Cf) NHWC -> (BL, H', W', d) NHWC."""
        BLn = feats[0].shape[0]
        return self.plan().run(feats, n, intrinsic.reshape(BLn, 3, 3).float().contiguous(),
                               extrinsic.reshape(BLn, 4, 4).float().contiguous(), invert_intrinsic=True)

    def forward(self, batch):
        b, l, n = batch['inputs'].shape[:3]
        feats = []
        for f in batch['features']:
            _require_inference(self, f)
            feats.append(ops.nchw_to_nhwc(f.reshape(b * l * n, *f.shape[3:]), self.compute_dtype))
        y = self.run_nhwc(feats, n, batch['intrinsic'], batch['extrinsic'])
        y = ops.nhwc_to_nchw(y, torch.float32)
        return y.view(b, l, *y.shape[1:])


# --------------------------------------------------------------------------- FuseBEVT
class PreNormResidual(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class SwapAttention(nn.Module):
    """swap_fusion_modules.Attention: parameters + the 3-D relative position index buffer."""

    def __init__(self, dim, dim_head=32, dropout=0., agent_size=6, window_size=7):
        super().__init__()
        assert (dim % dim_head) == 0, 'dimension should be divisible by dimension per head'
        assert dim_head == 32, "kernels are built for head_dim 32"
        self.heads = dim // dim_head
        self.scale = dim_head ** -0.5
        self.window_size = [agent_size, window_size, window_size]
        self.to_qkv = nn.Linear(dim, dim * 3, bias=False)
        self.attend = nn.Sequential(nn.Softmax(dim=-1))
        self.to_out = nn.Sequential(nn.Linear(dim, dim, bias=False), nn.Dropout(dropout))
        L, wh, ww = self.window_size
        self.relative_position_bias_table = nn.Embedding((2 * L - 1) * (2 * wh - 1) * (2 * ww - 1), self.heads)
        cd, ch, cw = torch.meshgrid(torch.arange(L), torch.arange(wh), torch.arange(ww), indexing='ij')
        c = torch.stack([cd.reshape(-1), ch.reshape(-1), cw.reshape(-1)], 0)
        rel = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0).contiguous()
        idx = ((rel[..., 0] + L - 1) * (2 * wh - 1) * (2 * ww - 1) + (rel[..., 1] + wh - 1) * (2 * ww - 1)
               + rel[..., 2] + ww - 1)
        self.register_buffer("relative_position_index", idx)


def _mask_u8(mask, B, H, W, L):
    """(B, H, W, 1, L) reference mask -> (B, H, W, L) uint8 key mask of the attention kernel."""
    if mask is None or tuple(mask.shape) != (B, H, W, 1, L):
        # the reference's rearrange 'b (x w1) (y w2) e l -> ...' (swap_fusion_modules.py:172-174) rejects other shapes
        raise _einops_error(f"mask must be (B, H, W, 1, L) = {(B, H, W, 1, L)}, got {None if mask is None else tuple(mask.shape)}")
    return (mask[:, :, :, 0, :] != 0).to(torch.uint8).contiguous()


class _SwapBlockBase(_Planned):
    """forward shared by SwapFusionBlockMask / SwapFusionBlock: x (B, L, d, H, W) -> same shape."""
    _use_mask = False

    def _build_plan(self, dev, dt):
        return engine.SwapBlockPlan(self, self._use_mask, self.window_size, dev, dt)

    def _run(self, x, mask):
        _require_inference(self, x)
        B, L, d, H, W = x.shape
        ws = self.window_size
        if H % ws or W % ws:
            raise _einops_error(f"map {H}x{W} is not divisible by window_size {ws}")      # :175 / Rearrange
        if L != self.agent_size:
            raise ValueError(f"agent axis {L} != agent_size {self.agent_size} baked into the bias table")
        t = ops.nchw_to_nhwc(x.reshape(B * L, d, H, W), self.compute_dtype).reshape(-1, d)
        mk = _mask_u8(mask, B, H, W, L) if self._use_mask else None
        y, _ = self.plan().run(t, None, (B, L, H, W, d), mk)
        return ops.nhwc_to_nchw(y.view(B * L, H, W, d), torch.float32).view(B, L, d, H, W)


class SwapFusionBlockMask(_SwapBlockBase):
    """swap_fusion_modules.py:131-192."""
    _use_mask = True

    def __init__(self, input_dim, mlp_dim, dim_head, window_size, agent_size, drop_out):
        super().__init__()
        self.window_size, self.agent_size = window_size, agent_size
        mk = lambda: SwapAttention(input_dim, dim_head, drop_out, agent_size, window_size)
        self.window_attention = PreNormResidual(input_dim, mk())
        self.window_ffd = PreNormResidual(input_dim, FeedForward(input_dim, mlp_dim, drop_out))
        self.grid_attention = PreNormResidual(input_dim, mk())
        self.grid_ffd = PreNormResidual(input_dim, FeedForward(input_dim, mlp_dim, drop_out))

    def forward(self, x, mask):
        return self._run(x, mask)


class SwapFusionBlock(_SwapBlockBase):
    """swap_fusion_modules.py:195-230 (``forward(x, mask=None)`` ignores the mask, :227-230)."""

    def __init__(self, input_dim, mlp_dim, dim_head, window_size, agent_size, drop_out):
        super().__init__()
        self.window_size, self.agent_size = window_size, agent_size
        mk = lambda: SwapAttention(input_dim, dim_head, drop_out, agent_size, window_size)
        # positions 0,3,4,7 are the reference's parameter-free Rearrange layers
        self.block = nn.Sequential(
            nn.Identity(), PreNormResidual(input_dim, mk()),
            PreNormResidual(input_dim, FeedForward(input_dim, mlp_dim, drop_out)), nn.Identity(),
            nn.Identity(), PreNormResidual(input_dim, mk()),
            PreNormResidual(input_dim, FeedForward(input_dim, mlp_dim, drop_out)), nn.Identity())

    def forward(self, x, mask=None):
        return self._run(x, None)


class SwapFusionEncoder(_Planned):
    def __init__(self, args):
        super().__init__()
        self.depth = args['depth']
        self.window_size = args['window_size']
        self.agent_size = args['agent_size']
        self.mask = args['mask'] if 'mask' in args else False
        cls = SwapFusionBlockMask if self.mask else SwapFusionBlock
        self.layers = nn.ModuleList([cls(args['input_dim'], args['mlp_dim'], args['dim_head'], args['window_size'],
                                         args['agent_size'], args['drop_out']) for _ in range(self.depth)])
        # indices 0,1,4 are the reference's Reduce / Rearrange layers
        self.mlp_head = nn.Sequential(nn.Identity(), nn.Identity(), nn.LayerNorm(args['input_dim']),
                                      nn.Linear(args['input_dim'], args['input_dim']), nn.Identity())

    def _build_plan(self, dev, dt):
        return engine.FusionPlan(self, dev, dt)

    def run_nhwc(self, x, mask_u8):
        B, L, H, W, _ = x.shape
        if L != self.agent_size:
            raise ValueError(f"agent axis {L} != agent_size {self.agent_size} baked into the bias table")
        if H % self.window_size or W % self.window_size:
            raise _einops_error(f"map {H}x{W} is not divisible by window_size {self.window_size}")
        return self.plan().run(x, mask_u8)

    def forward(self, x, mask=None):
        """x (B, L, d, H, W); mask (B, H, W, 1, L) -> (B, d, H, W)."""
        _require_inference(self, x)
        B, L, d, H, W = x.shape
        t = ops.nchw_to_nhwc(x.reshape(B * L, d, H, W), self.compute_dtype).view(B, L, H, W, d)
        mk = _mask_u8(mask, B, H, W, L) if self.mask else None
        y = self.run_nhwc(t, mk)
        return ops.nhwc_to_nchw(y, torch.float32)


# --------------------------------------------------------------------------- decoder / head
class NaiveDecoder(_Planned):
    def __init__(self, params):
        super().__init__()
        self.num_ch_dec, self.num_layer, self.input_dim = params['num_ch_dec'], params['num_layer'], params['input_dim']
        assert len(self.num_ch_dec) == self.num_layer
        mods = []
        for i in range(self.num_layer - 1, -1, -1):
            cin = self.input_dim if i == self.num_layer - 1 else self.num_ch_dec[i + 1]
            cout = self.num_ch_dec[i]
            mods += [nn.Conv2d(cin, cout, 3, 1, 1), nn.BatchNorm2d(cout), nn.ReLU(True),
                     nn.Conv2d(cout, cout, 3, 1, 1), nn.BatchNorm2d(cout), nn.ReLU(True)]
        self.decoder = nn.ModuleList(mods)

    def _build_plan(self, dev, dt):
        return engine.DecoderPlan(self, dev, dt)

    def forward(self, x):
        _require_inference(self, x)
        b, l, c, h, w = x.shape
        y = self.plan().run(ops.nchw_to_nhwc(x.reshape(b * l, c, h, w), self.compute_dtype))
        y = ops.nhwc_to_nchw(y, torch.float32)
        return y.view(b, l, *y.shape[1:])


class BevSegHead(_Planned):
    def __init__(self, target, input_dim, output_class):
        super().__init__()
        self.target = target
        if target == 'dynamic':
            self.dynamic_head = nn.Conv2d(input_dim, output_class, 3, padding=1)
        if target == 'static':
            self.static_head = nn.Conv2d(input_dim, output_class, 3, padding=1)
        else:   # reference quirk: a non-'static' target also builds both heads (bev_seg_head.py:19-33)
            self.dynamic_head = nn.Conv2d(input_dim, output_class, 3, padding=1)
            self.static_head = nn.Conv2d(input_dim, output_class, 3, padding=1)

    def _build_plan(self, dev, dt):
        return engine.SegHeadPlan(self, dev, dt)

    def forward(self, x, b, l):
        _require_inference(self, x)
        return self.plan().run(ops.nchw_to_nhwc(x, self.compute_dtype), b, l)


# --------------------------------------------------------------------------- whole models
class STTF(nn.Module):
    """corpbevt.py:22-64: affine warp of every agent's BEV map into the ego frame."""

    def __init__(self, args):
        super().__init__()
        self.discrete_ratio = args['resolution']
        self.downsample_rate = args['downsample_rate']
        self.compute_dtype = COMPUTE_DTYPE

    def forward(self, x, spatial_correction_matrix):
        """x (B, L, C, H, W), spatial_correction_matrix (B, L, 4, 4) -> (B, L, H, W, C)."""
        _require_inference(self, x)
        B, L, C, H, W = x.shape
        t = ops.nchw_to_nhwc(x.reshape(B * L, C, H, W), self.compute_dtype)
        tm = spatial_correction_matrix.to(device=x.device, dtype=torch.float32).contiguous()
        full = torch.full((B,), L, dtype=torch.int64, device=x.device)
        y, _ = ops.sttf_regroup(t, tm, full, B, L, self.discrete_ratio, self.downsample_rate, False)
        return y.float()


class NaiveCompressor(_Planned):
    """naive_compress.py:5-28 (``compress_raito`` is the reference's spelling)."""

    def __init__(self, input_dim, compress_raito):
        super().__init__()
        mid = input_dim // compress_raito
        self.encoder = nn.Sequential(nn.Conv2d(input_dim, mid, 3, 1, 1), nn.BatchNorm2d(mid, eps=1e-3, momentum=0.01), nn.ReLU())
        self.decoder = nn.Sequential(nn.Conv2d(mid, input_dim, 3, 1, 1), nn.BatchNorm2d(input_dim, eps=1e-3, momentum=0.01), nn.ReLU(),
                                     nn.Conv2d(input_dim, input_dim, 3, 1, 1), nn.BatchNorm2d(input_dim, eps=1e-3, momentum=0.01),
                                     nn.ReLU())

    def _build_plan(self, dev, dt):
        return engine.CompressorPlan(self, dev, dt)

    def forward(self, x):
        _require_inference(self, x)
        return ops.nhwc_to_nchw(self.plan().run(ops.nchw_to_nhwc(x, self.compute_dtype)), torch.float32)


def post_process(output_dict):
    """CameraBevPostprocessor.post_process_train (opv2v/opencood/data_utils/post_processor/
    camera_bev_postprocessor.py:55-94): per head softmax over classes of seg[:, 0] and its argmax, one kernel."""
    out = dict(output_dict)
    for key in ("static", "dynamic"):
        prob, amap = ops.softmax_argmax_nchw(output_dict[key + "_seg"][:, 0])
        out[key + "_prob"], out[key + "_map"] = prob, amap
    return out


class CorpBEVT(nn.Module):
    """corpbevt.py:67-145.  forward(batch_dict) -> {'static_seg','dynamic_seg'} (B,1,classes,H,W).

    ``write_features`` (default False): also replicate the reference's side effect
    ``batch_dict['features'] = encoder(x)`` (:113) -- it costs a NHWC->NCHW fp32 copy of every picked level."""

    def __init__(self, config):
        super().__init__()
        self.max_cav = config['max_cav']
        self.encoder = ResnetEncoder(config['encoder'])
        fax_params = config['fax']
        fax_params['backbone_output_shape'] = self.encoder.output_shapes
        self.fax = FAXModule(fax_params)
        if config['compression'] > 0:
            self.compression = True
            self.naive_compressor = NaiveCompressor(128, config['compression'])
        else:
            self.compression = False
        self.downsample_rate = config['sttf']['downsample_rate']
        self.discrete_ratio = config['sttf']['resolution']
        self.use_roi_mask = config['sttf']['use_roi_mask']
        self.sttf = STTF(config['sttf'])
        self.fusion_net = SwapFusionEncoder(config['fax_fusion'])
        if self.fusion_net.mask and not self.use_roi_mask:
            # survey quirk Q8: the (B,1,1,1,L) mask cannot be partitioned (swap_fusion_modules.py:172-174)
            raise _einops_error("mask=True requires use_roi_mask=True")
        self.decoder = NaiveDecoder(config['decoder'])
        self.target = config['target']
        self.seg_head = BevSegHead(self.target, config['seg_head_dim'], config['output_class'])
        self.write_features = False

    def encode_nhwc(self, inputs, intrinsic, extrinsic, batch_dict=None):
        """encoder -> SinBEVT (-> compressor): (rows, 1, M, H, W, 3) -> (rows, h, w, 128) NHWC 16-bit."""
        m = inputs.shape[2]
        feats = self.encoder.features_nhwc(inputs)
        pick = self.encoder.idx_pick
        feats = [feats[i] for i in pick]
        if self.write_features and batch_dict is not None:
            b, l = inputs.shape[:2]
            batch_dict.update({'features': [ops.nhwc_to_nchw(f, torch.float32).view(b, l, m, f.shape[3], f.shape[1], f.shape[2])
                                            for f in feats]})
        y = self.fax.run_nhwc(feats, m, intrinsic, extrinsic)                                 # (b*l, H, W, d)
        if self.compression:
            y = self.naive_compressor.plan().run(y)
        return y

    def fuse_nhwc(self, y, tm, record_len):
        """regroup/STTF/mask -> FuseBEVT -> decoder -> head on device-resident tm (B,L,4,4) fp32, record_len (B) int64."""
        B = record_len.shape[0]
        y, mask = ops.sttf_regroup(y, tm, record_len, B, self.max_cav, self.discrete_ratio, self.downsample_rate,
                                   self.use_roi_mask)
        y = self.fusion_net.run_nhwc(y, mask)                                                # (B, H, W, d)
        y = self.decoder.plan().run(y)
        return self.seg_head.plan().run(y, B, 1)

    def forward(self, batch_dict):
        x = batch_dict['inputs']
        y = self.encode_nhwc(x, batch_dict['intrinsic'], batch_dict['extrinsic'], batch_dict)
        record_len = batch_dict['record_len'].to(device=y.device, dtype=torch.int64)
        tm = batch_dict['transformation_matrix'].to(device=y.device, dtype=torch.float32).contiguous()
        return self.fuse_nhwc(y, tm, record_len)


class FaxFusedTransformer(nn.Module):
    """fax_fused_transformer.py:13-48 (SinBEVT only: encoder -> FAX -> decoder -> head)."""

    def __init__(self, config):
        super().__init__()
        self.encoder = ResnetEncoder(config['encoder'])
        cvm = config['fax']
        cvm['backbone_output_shape'] = self.encoder.output_shapes
        self.fax = FAXModule(cvm)
        self.decoder = NaiveDecoder(config['decoder'])
        self.target = config['target']
        self.seg_head = BevSegHead(self.target, config['seg_head_dim'], config['output_class'])

    def forward(self, batch_dict):
        x = batch_dict['inputs']
        b, l, m = x.shape[:3]
        feats = self.encoder.features_nhwc(x)
        feats = [feats[i] for i in self.encoder.idx_pick]
        y = self.fax.run_nhwc(feats, m, batch_dict['intrinsic'], batch_dict['extrinsic'])
        y = self.decoder.plan().run(y)
        return self.seg_head.plan().run(y, b, l)
