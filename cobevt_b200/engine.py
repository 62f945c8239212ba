"""Inference engine of the FAX hot path: weight plans (folded + packed once) and the kernel
sequences that run them.  Everything between the model's input and output tensors stays
NHWC 16-bit in HBM; every arithmetic step is a call into libcobevt_sm100.so (ops.py).

Folding done at plan time (all exact re-associations of the reference math):
  * eval BatchNorm after a conv          -> scaled weights + bias
  * LayerNorm affine in front of a Linear -> W diag(gamma), W beta + b   (kernels then only
    ever need the un-affined normalised row, produced by the preceding epilogue)
  * PixelUnshuffle(2) between two convs   -> a 6x3 stride-(2,1) conv over the (H, W/2, 2C) view
  * nearest x2 upsample before a conv     -> index shift in the implicit-GEMM gather
  * attn2's `repeat(query, n)` + mean     -> one copy (survey quirk Q1)
"""
import torch

from . import ops
from .ops import PackedConv, PackedMlp, ACT_NONE, ACT_RELU, ACT_GELU, ACT_SWISH

LN_EPS = 1e-5


# --------------------------------------------------------------------------- folding helpers
def _bn_scale_shift(bn):
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return s, bn.bias.detach().float() - bn.running_mean.detach().float() * s


def fold_conv_bn(conv, bn):
    """conv (no act) followed by eval BatchNorm -> (weight, bias)."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0])
    if bn is None:
        return w, (b if conv.bias is not None else None)
    s, t = _bn_scale_shift(bn)
    s, t = s.to(w.device), t.to(w.device)
    return w * s.view(-1, 1, 1, 1), b.to(w.device) * s + t


def fold_ln_linear(ln, lin):
    """Linear(LayerNorm(x)) = (W diag(g)) xhat + (W beta + b)."""
    w = lin.weight.detach().float()
    g, be = ln.weight.detach().float(), ln.bias.detach().float()
    b = w @ be
    if lin.bias is not None:
        b = b + lin.bias.detach().float()
    return w * g[None, :], b


def _dev_f32(t, device):
    return t.detach().float().contiguous().to(device)


# --------------------------------------------------------------------------- ResNet trunk
class ResNetPlan:
    """BasicBlock / Bottleneck ResNet trunk up to the deepest picked level
    (resnet_ms.py:46-89; levels past max(id_pick) cannot influence the output and are skipped)."""

    def __init__(self, trunk, device, dtype, n_levels):
        self.dtype = dtype
        w, b = fold_conv_bn(trunk.conv1, trunk.bn1)
        # 7x7/s2 stem on a zero-bordered RGBx (4-channel) image viewed as "super-pixels" of 2 pixels x 4
        # channels = 16 bytes: (N, H+6, Wp/2, 8).  A kernel row is 4 super-taps (8 pixels, the 8th has zero
        # weight), stride (2,1), no padding -> K = 7*4*8 = 224 (vs 392 for RGB padded to 8 channels), no
        # bounds checks, every tap 16-byte aligned.
        co = w.shape[0]
        wp = torch.zeros(co, 4, 7, 8)
        wp[:, :3, :, :7] = w                                    # (co, c, ky, kx) with c<3, kx<7
        wp = wp.reshape(co, 4, 7, 4, 2).permute(0, 4, 1, 2, 3)   # co dx c ky kxs
        self.stem = PackedConv(wp.reshape(co, 8, 7, 4), b, device, dtype)
        self.stem.k_alg = 7 * 7 * 3
        self.levels = []
        for li in range(n_levels):
            blocks = []
            for blk in getattr(trunk, f"layer{li + 1}"):
                convs = []
                names = [("conv1", "bn1"), ("conv2", "bn2")] + ([("conv3", "bn3")] if hasattr(blk, "conv3") else [])
                for cn, bnn in names:
                    conv = getattr(blk, cn)
                    w, b = fold_conv_bn(conv, getattr(blk, bnn))
                    convs.append((PackedConv(w, b, device, dtype), conv.stride[0], conv.padding[0]))
                ds = None
                if blk.downsample is not None:
                    w, b = fold_conv_bn(blk.downsample[0], blk.downsample[1])
                    ds = (PackedConv(w, b, device, dtype), blk.downsample[0].stride[0])
                blocks.append((convs, ds))
            self.levels.append(blocks)

    def run(self, images, nchw=False, mean=None, std=None, bgr=False):
        """images (NI, H, W, 3) [or (NI, 3, H, W) with nchw] fp32, or uint8 camera frames (then x/255, optional
        BGR->RGB and (x - mean) / std happen in the staging kernel) -> list of NHWC 16-bit feature maps."""
        if nchw:
            N, _, H, W = images.shape
        else:
            N, H, W, _ = images.shape
        OH, OW = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        Hp, Wp = max(H + 6, 2 * OH + 5), (max(W + 6, 2 * OW + 6) + 1) // 2 * 2
        x = ops.pack_rgb4_padded(images, self.dtype, 3, 3, Hp, Wp, nchw=nchw, mean=mean, std=std, bgr=bgr).view(N, Hp, Wp // 2, 8)
        x, _, _ = ops.conv2d_nhwc(x, self.stem, stride=(2, 1), pad=0, act=ACT_RELU, out_hw=(OH, OW))
        x = ops.maxpool3x3s2(x)
        feats = []
        for blocks in self.levels:
            for convs, ds in blocks:
                idt = x
                if ds is not None:
                    idt, _, _ = ops.conv2d_nhwc(x, ds[0], stride=ds[1], pad=0)
                y = x
                for ci, (pc, st, pd) in enumerate(convs):
                    last = ci == len(convs) - 1
                    y, _, _ = ops.conv2d_nhwc(y, pc, stride=st, pad=pd, act=ACT_RELU, res=idt if last else None)
                x = y
            feats.append(x)
        return feats


class BottleneckPlan:
    """torchvision Bottleneck(c, c//4) with eval BN folded (fax_modules.py:10,472)."""

    def __init__(self, blk, device, dtype):
        self.c = [PackedConv(*fold_conv_bn(getattr(blk, f"conv{i}"), getattr(blk, f"bn{i}")), device, dtype) for i in (1, 2, 3)]

    def run(self, x):
        y, _, _ = ops.conv2d_nhwc(x, self.c[0], act=ACT_RELU)
        y, _, _ = ops.conv2d_nhwc(y, self.c[1], pad=1, act=ACT_RELU)
        y, _, _ = ops.conv2d_nhwc(y, self.c[2], act=ACT_RELU, res=x)
        return y


class DownsamplePlan:
    """conv3x3 -> PixelUnshuffle(2) -> conv3x3-BN-ReLU -> conv1x1-BN (fax_modules.py:476-489)."""

    def __init__(self, seq, device, dtype):
        conv_a, _, conv_b, bn_b, _, conv_c, bn_c = list(seq)
        self.a = PackedConv(conv_a.weight.detach().float(), None, device, dtype)
        cq = conv_a.weight.shape[0]                      # channels before the unshuffle
        wb, bb = fold_conv_bn(conv_b, bn_b)              # (Co, cq*4, 3, 3), in-channel = c*4 + dy*2 + dx
        co = wb.shape[0]
        wb = wb.reshape(co, cq, 2, 2, 3, 3)              # co c dy dx ky kx
        wb = wb.permute(0, 3, 1, 4, 2, 5)                # co dx c ky dy kx
        wb = wb.reshape(co, 2 * cq, 6, 3)                # in-channel = dx*cq + c ; kh = 2*ky + dy
        self.b = PackedConv(wb, bb, device, dtype)
        self.c = PackedConv(*fold_conv_bn(conv_c, bn_c), device, dtype)
        self.cq = cq

    def run(self, x, want_ln):
        N, H, W, _ = x.shape
        y, _, _ = ops.conv2d_nhwc(x, self.a, pad=1)                       # (N, H, W, cq)
        y = y.view(N, H, W // 2, 2 * self.cq)                            # merge dx into channels
        y, _, _ = ops.conv2d_nhwc(y, self.b, stride=(2, 1), pad=(2, 1), act=ACT_RELU)
        return ops.conv2d_nhwc(y, self.c, ln_out=want_ln)                # x, None, LN0(x)


# --------------------------------------------------------------------------- FAX stage
class CrossAttnPlan:
    """CrossWinAttention weights: LayerNorm affines folded into q/k/v projections."""

    def __init__(self, m, device, dtype):
        self.q = PackedConv(*fold_ln_linear(m.to_q[0], m.to_q[1]), device, dtype)
        wk, bk = fold_ln_linear(m.to_k[0], m.to_k[1])
        wv, bv = fold_ln_linear(m.to_v[0], m.to_v[1])
        self.wk, self.bk, self.wv, self.bv = wk, bk, wv, bv      # kept for the fused two-attention projection
        self.k = PackedConv(wk, bk, device, dtype)
        self.v = PackedConv(wv, bv, device, dtype)
        # a zero-padded key/value row becomes Linear(LayerNorm(0)) = W beta + b (quirk Q2)
        self.k_pad = bk.to(dtype).to(device)
        self.v_pad = bv.to(dtype).to(device)
        self.proj = PackedConv(m.proj.weight.detach().float(), m.proj.bias.detach().float(), device, dtype)
        self.heads, self.dim_head = m.heads, m.dim_head
        self.scale = m.scale


class FaxStagePlan:
    def __init__(self, cva, world, device, dtype):
        self.dtype = dtype
        self.dim = cva.prenorm_1.weight.shape[0]
        self.q_win, self.feat_win = tuple(cva.q_win_size), tuple(cva.feat_win_size)
        self.bev_flag, self.skip = cva.bev_embed_flag, cva.skip
        self.image_h, self.image_w = cva.image_height, cva.image_width
        self.no_img_feat = cva.feature_proj is None
        sl, tl = _bn_scale_shift(cva.feature_linear[0])
        self.bn_l = (_dev_f32(sl, device), _dev_f32(tl, device))
        self.f_lin = PackedConv(cva.feature_linear[2].weight.detach().float(), None, device, dtype)
        if not self.no_img_feat:
            sp, tp = _bn_scale_shift(cva.feature_proj[0])
            self.bn_p = (_dev_f32(sp, device), _dev_f32(tp, device))
            self.f_proj = PackedConv(cva.feature_proj[2].weight.detach().float(), None, device, dtype)
        self.w_img = _dev_f32(cva.img_embed.weight.reshape(self.dim, 4), device)
        self.w_cam = _dev_f32(cva.cam_embed.weight.reshape(self.dim, 4), device)
        if self.bev_flag:
            self.w_bev = _dev_f32(cva.bev_embed.weight.reshape(self.dim, 2), device)
            self.b_bev = _dev_f32(cva.bev_embed.bias, device)
            self.world = _dev_f32(world.reshape(2, -1), device)
        self.a1 = CrossAttnPlan(cva.cross_win_attend_1, device, dtype)
        self.a2 = CrossAttnPlan(cva.cross_win_attend_2, device, dtype)
        if self.bev_flag:
            # analytic query projection (csrc/eltwise2.cu bev_q_kernel): q = W' LN0(x + e/|e|) + b' with e affine in the
            # cell's world coordinates -- per-channel constants of W' applied to the embedding vectors, fp32 on the host
            m = cva.cross_win_attend_1
            wq, bq = fold_ln_linear(m.to_q[0], m.to_q[1])
            wq, bq = wq.cpu(), bq.cpu()
            wb = cva.bev_embed.weight.detach().float().reshape(self.dim, 2).cpu()
            bb = cva.bev_embed.bias.detach().float().cpu()
            wc = cva.cam_embed.weight.detach().float().reshape(self.dim, 4).cpu()
            a_, b_ = wb[:, 0], wb[:, 1]
            z = torch.zeros_like(a_)
            tab = torch.stack([a_, b_, bb] + [wc[:, k] for k in range(4)] + [z] + [wq @ a_, wq @ b_, wq.sum(1), bq] +
                              [wq @ bb] + [wq @ wc[:, k] for k in range(4)] + [z, z, z], dim=1)          # (C, 20)
            self.q_tab = tab.contiguous().to(device)
            self.q_consts = (float(a_ @ a_), float(b_ @ b_), float(2 * (a_ @ b_)), float(a_.mean()), float(b_.mean()))
            self.q_wx = PackedConv(wq, None, device, dtype)
        # both attentions project the same normalised key / value maps: one N=256 GEMM each
        self.k12 = PackedConv(torch.cat([self.a1.wk, self.a2.wk], 0), torch.cat([self.a1.bk, self.a2.bk], 0), device, dtype)
        self.v12 = PackedConv(torch.cat([self.a1.wv, self.a2.wv], 0), torch.cat([self.a1.bv, self.a2.bv], 0), device, dtype)
        self.m1 = PackedMlp(*fold_ln_linear(cva.prenorm_1, cva.mlp_1[0]), cva.mlp_1[2].weight.detach().float(),
                            cva.mlp_1[2].bias.detach().float(), device, dtype)
        self.m2 = PackedMlp(*fold_ln_linear(cva.prenorm_2, cva.mlp_2[0]), cva.mlp_2[2].weight.detach().float(),
                            cva.mlp_2[2].bias.detach().float(), device, dtype)
        self.post_g = _dev_f32(cva.postnorm.weight, device)
        self.post_b = _dev_f32(cva.postnorm.bias, device)

    def run(self, x, x_ln, x_rows, feat, n, intrinsic, E_inv, invert_intrinsic=True):
        """x: (rows, H, W, d) residual stream (rows == BL, or 1 for the broadcast prior);
        x_ln: LayerNorm0(x) or None; feat: (BL*n, h, w, Cf).  Returns (BL, H, W, d)."""
        d = self.dim
        BLn, h, w, _ = feat.shape
        BL = BLn // n
        H, W = x.shape[1:3]
        P = H * W
        # ---- key / value token maps (fax_modules.py:343-358,377-384) -------------------------
        if self.no_img_feat:
            fb, _ = ops.bn_relu2(feat, *self.bn_l)
        else:
            fa, fb = ops.bn_relu2(feat, *self.bn_p, *self.bn_l)
        img = ops.ray_embed(intrinsic, E_inv, self.w_img, self.w_cam, h, w, self.image_h, self.image_w, d,
                            self.dtype, invert_intrinsic=invert_intrinsic)
        if self.no_img_feat:
            k_hat = ops.layernorm_rows(img)
        else:
            _, _, k_hat = ops.conv2d_nhwc(fa, self.f_proj, res=img, want16=False, ln_out=True)
        _, _, v_hat = ops.conv2d_nhwc(fb, self.f_lin, want16=False, ln_out=True)
        k12 = ops.conv2d_nhwc(k_hat, self.k12)[0].view(BL, n, h, w, 2 * d)
        v12 = ops.conv2d_nhwc(v_hat, self.v12)[0].view(BL, n, h, w, 2 * d)
        k1, k2 = k12[..., :d], k12[..., d:]            # column slices: token stride 2d, no copies
        v1, v2 = v12[..., :d], v12[..., d:]
        fh, fw = self.feat_win
        hp = h if h % fh == 0 else ((h + fh) // fh) * fh                   # pad_divisble, :315-321
        wp = w if w % fw == 0 else ((w + fw) // fw) * fw
        padded = (hp, wp) != (h, w)
        # ---- queries ---------------------------------------------------------------------------
        if self.bev_flag and d in (32, 64, 128) and n <= min(8, d // 8):
            q1 = ops.bev_q_analytic(x.reshape(-1, d), self.q_wx, self.q_tab, self.q_consts, self.world, E_inv, BL, n, P)
            q1 = q1.view(BL, n, H, W, d)
            nq = n
        else:
            if self.bev_flag:
                q_hat = ops.bev_query(x.reshape(-1, d), self.world, self.w_bev, self.b_bev, self.w_cam, E_inv, BL, n, P)
                nq = n
            else:
                q_hat = x_ln if x_ln is not None else ops.layernorm_rows(x)
                nq = 1
            q1 = ops.linear_tokens(q_hat.reshape(-1, d), self.a1.q)[0].view(BL, nq, H, W, d)
        qm = ops.token_map(nq, H, W, *self.q_win, False)
        km = ops.token_map(n, hp, wp, fh, fw, False, Hv=h, Wv=w)
        a = ops.window_attention(q1, k1, v1, qm, km, self.a1.heads, self.a1.scale, True,
                                 k_pad=self.a1.k_pad if padded else None, v_pad=self.a1.v_pad if padded else None)
        skip = x.reshape(-1, d) if self.skip else None
        rr = P * x_rows if (self.skip and x_rows != BL) else 0
        z, _, z_ln = ops.linear_tokens(a.reshape(-1, d), self.a1.proj, res=skip, res_rows=rr, ln_out=True)
        z, z_ln = ops.token_mlp(z_ln, self.m1, z)
        # ---- local window <-> dilated grid (fax_modules.py:413-433) ---------------------------
        q2 = ops.linear_tokens(z_ln, self.a2.q)[0].view(BL, 1, H, W, d)
        qm = ops.token_map(1, H, W, *self.q_win, False)
        km = ops.token_map(n, hp, wp, fh, fw, True, Hv=h, Wv=w)
        a = ops.window_attention(q2, k2, v2, qm, km, self.a2.heads, self.a2.scale, True,
                                 k_pad=self.a2.k_pad if padded else None, v_pad=self.a2.v_pad if padded else None)
        z2, _, z_ln = ops.linear_tokens(a.reshape(-1, d), self.a2.proj, res=z if self.skip else None, ln_out=True)
        _, out = ops.token_mlp(z_ln, self.m2, z2, want16=False, ln_out=True, ln_gamma=self.post_g, ln_beta=self.post_b)
        return out.view(BL, H, W, d)


class GlobalAttnPlan:
    """fax_modules.Attention (:93-176): 1024-token self-attention with 2-D rel-pos bias."""

    def __init__(self, m, device, dtype):
        self.qkv = PackedConv(m.to_qkv.weight.detach().float(), None, device, dtype)
        self.out = PackedConv(m.to_out[0].weight.detach().float(), None, device, dtype)
        self.heads, self.scale = m.heads, m.scale
        bias = m.rel_pos_bias.weight.detach().float()[m.rel_pos_indices.to(m.rel_pos_bias.weight.device)]  # T T h
        self.bias = ops.AttnBias(bias.permute(2, 0, 1).contiguous().to(device))

    def run(self, x):
        BL, H, W, d = x.shape
        qkv = ops.linear_tokens(x, self.qkv)[0].view(BL, 1, H, W, 3 * d)
        m = ops.token_map(1, H, W, H, W, False)
        a = ops.window_attention(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], m, m, self.heads, self.scale,
                                 False, bias=self.bias)
        return ops.linear_tokens(a.view(BL, H, W, d), self.out)[0]


class FaxPlan:
    def __init__(self, fax, device, dtype):
        self.dtype = dtype
        be = fax.bev_embedding
        self.stages = [FaxStagePlan(cv, getattr(be, f"grid{i}")[:2], device, dtype)
                       for i, cv in enumerate(fax.cross_views)]
        self.layers = [[BottleneckPlan(b, device, dtype) for b in seq] for seq in fax.layers]
        self.down = [DownsamplePlan(seq[0], device, dtype) for seq in fax.downsample_layers]
        sa = getattr(fax, 'self_attn', None)
        self.self_attn = GlobalAttnPlan(sa, device, dtype) if sa is not None else None
        prior = be.learned_features.detach().float().permute(1, 2, 0).contiguous()         # H W d
        self.prior = prior.to(dtype).to(device)[None]

    def run(self, feats, n, intrinsic, E_inv, invert_intrinsic=True):
        """feats: list of (BL*n, h, w, Cf) NHWC; intrinsic (BL*n,3,3), E_inv (BL*n,4,4) fp32."""
        BL = feats[0].shape[0] // n
        x, x_ln, x_rows = self.prior, None, 1
        for i, st in enumerate(self.stages):
            x = st.run(x, x_ln, x_rows, feats[i], n, intrinsic, E_inv, invert_intrinsic)
            x_rows = BL
            for b in self.layers[i]:
                x = b.run(x)
            x_ln = None
            if i < len(self.stages) - 1:
                nxt_needs_ln = not self.stages[i + 1].bev_flag
                x, _, x_ln = self.down[i].run(x, nxt_needs_ln)
        if self.self_attn is not None:
            x = self.self_attn.run(x)
        return x


# --------------------------------------------------------------------------- FuseBEVT
class SwapAttnPlan:
    def __init__(self, prenorm_attn, prenorm_ffd, device, dtype):
        at = prenorm_attn.fn
        self.qkv = PackedConv(*fold_ln_linear(prenorm_attn.norm, at.to_qkv), device, dtype)
        self.out = PackedConv(at.to_out[0].weight.detach().float(), None, device, dtype)
        tbl = at.relative_position_bias_table.weight.detach().float()
        idx = at.relative_position_index.to(tbl.device)
        self.bias = ops.AttnBias(tbl[idx].permute(2, 0, 1).contiguous().to(device))      # heads T T
        self.heads, self.scale = at.heads, at.scale
        ff = prenorm_ffd.fn.net
        self.ff = PackedMlp(*fold_ln_linear(prenorm_ffd.norm, ff[0]), ff[3].weight.detach().float(),
                            ff[3].bias.detach().float(), device, dtype)


class SwapBlockPlan:
    """One SwapFusionBlockMask / SwapFusionBlock (swap_fusion_modules.py:131-230): window attention + FFN, then
    grid attention + FFN, both PreNormResidual; the partitions are addressing inside the attention kernel."""

    def __init__(self, blk, use_mask, ws, device, dtype):
        self.use_mask, self.ws = use_mask, ws
        if use_mask:
            pairs = [(blk.window_attention, blk.window_ffd), (blk.grid_attention, blk.grid_ffd)]
        else:
            pairs = [(blk.block[1], blk.block[2]), (blk.block[5], blk.block[6])]
        self.pairs = [SwapAttnPlan(a, f, device, dtype) for a, f in pairs]

    def run(self, x, x_ln, dims, mask):
        """x (B*L*H*W, d) residual stream, x_ln its plain LayerNorm (or None) -> (x, x_ln) after the block."""
        B, L, H, W, d = dims
        if x_ln is None:
            x_ln = ops.layernorm_rows(x)
        for kind, sp in enumerate(self.pairs):
            qkv = ops.linear_tokens(x_ln, sp.qkv)[0].view(B, L, H, W, 3 * d)
            m = ops.token_map(L, H, W, self.ws, self.ws, kind == 1)
            a = ops.window_attention(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], m, m, sp.heads, sp.scale,
                                     False, bias=sp.bias, key_mask=mask if self.use_mask else None)
            x, _, x_ln = ops.linear_tokens(a.reshape(-1, d), sp.out, res=x, ln_out=True)
            x, x_ln = ops.token_mlp(x_ln, sp.ff, x)
        return x, x_ln


class FusionPlan:
    """SwapFusionEncoder (swap_fusion_modules.py:233-286)."""

    def __init__(self, enc, device, dtype):
        self.use_mask = enc.mask
        self.ws = enc.window_size
        self.blocks = [SwapBlockPlan(blk, enc.mask, enc.window_size, device, dtype) for blk in enc.layers]
        self.head = PackedConv(*fold_ln_linear(enc.mlp_head[2], enc.mlp_head[3]), device, dtype)

    def run(self, x, mask):
        """x (B, L, H, W, d) 16-bit; mask (B, H, W, L) uint8 or None -> (B, H, W, d)."""
        B, L, H, W, d = x.shape
        x = x.reshape(-1, d)
        x_ln = None
        for blk in self.blocks:
            x, x_ln = blk.run(x, x_ln, (B, L, H, W, d), mask)
        m_ln = ops.mean_views_ln(x.view(B, L, H * W, d))
        return ops.linear_tokens(m_ln, self.head)[0].view(B, H, W, d)


class CompressorPlan:
    """NaiveCompressor (naive_compress.py:5-28): three conv3x3 + BatchNorm(eps 1e-3) + ReLU."""

    def __init__(self, m, device, dtype):
        self.convs = [PackedConv(*fold_conv_bn(c, b), device, dtype)
                      for c, b in ((m.encoder[0], m.encoder[1]), (m.decoder[0], m.decoder[1]), (m.decoder[3], m.decoder[4]))]

    def run(self, x):
        for pc in self.convs:
            x, _, _ = ops.conv2d_nhwc(x, pc, pad=1, act=ACT_RELU)
        return x


# --------------------------------------------------------------------------- decoder / head
class DecoderPlan:
    """NaiveDecoder (naive_decoder.py:62-91): per level conv-BN-ReLU, nearest x2, conv-BN-ReLU."""

    def __init__(self, dec, device, dtype):
        mods = list(dec.decoder)
        self.levels = []
        for j in range(len(mods) // 6):
            a = PackedConv(*fold_conv_bn(mods[6 * j], mods[6 * j + 1]), device, dtype)
            b = PackedConv(*fold_conv_bn(mods[6 * j + 3], mods[6 * j + 4]), device, dtype)
            self.levels.append((a, b))

    def run(self, x):
        for a, b in self.levels:
            x, _, _ = ops.conv2d_nhwc(x, a, pad=1, act=ACT_RELU)
            x, _, _ = ops.conv2d_nhwc(x, b, pad=1, act=ACT_RELU, in_upsample=True)
        return x


class SegHeadPlan:
    def __init__(self, head, device, dtype):
        self.target = head.target
        mk = lambda c: PackedConv(c.weight.detach().float(), c.bias.detach().float(), device, dtype)
        # bev_seg_head.py:35-61: 'dynamic' -> dynamic head only, 'static' -> static head only, anything else -> both
        self.dyn = mk(head.dynamic_head) if head.target != "static" else None
        self.sta = mk(head.static_head) if head.target != "dynamic" else None

    def run(self, x, b, l):
        """x (b*l, H, W, C) -> dict of (b, l, classes, H, W) fp32 (bev_seg_head.py:35-61)."""
        out = {}
        for key, pc in (("dynamic_seg", self.dyn), ("static_seg", self.sta)):
            if pc is not None:
                _, y32, _ = ops.conv2d_nhwc(x, pc, pad=1, want16=False, out32=True)
                y = ops.nhwc_to_nchw(y32, torch.float32)
                out[key] = y.view(b, l, *y.shape[1:])
        if "dynamic_seg" not in out:
            out["dynamic_seg"] = torch.zeros_like(out["static_seg"])
        if "static_seg" not in out:
            out["static_seg"] = torch.zeros_like(out["dynamic_seg"])
        return out


# --------------------------------------------------------------------------- nuScenes SinBEVT (CVT fork)
class PyramidPlan:
    """PyramidAxialEncoder (nuscenes/cross_view_transformer/model/encoder_pyramid_axial.py:475-558):
    the same FAX stages / Bottlenecks / downsample blocks with per-stage widths (32/64/128, heads
    1/2/4), no global self-attention, intrinsics AND extrinsics inverted (:538-539)."""

    def __init__(self, enc, device, dtype):
        self.dtype = dtype
        be = enc.bev_embedding
        self.stages = [FaxStagePlan(cv, getattr(be, f"grid{i}")[:2], device, dtype) for i, cv in enumerate(enc.cross_views)]
        self.layers = [[BottleneckPlan(b, device, dtype) for b in seq] for seq in enc.layers]
        self.down = [DownsamplePlan(seq[0], device, dtype) for seq in enc.downsample_layers]
        prior = be.learned_features.detach().float().permute(1, 2, 0).contiguous()
        self.prior = prior.to(dtype).to(device)[None]

    def run(self, feats, n, intrinsics, extrinsics):
        """feats: list of (b*n, h, w, Cf) NHWC; intrinsics (b*n,3,3), extrinsics (b*n,4,4) fp32 (not inverted)."""
        b = feats[0].shape[0] // n
        E_inv = ops.invert_matrices(extrinsics)
        I_inv = ops.invert_matrices(intrinsics)
        x, x_ln, x_rows = self.prior, None, 1
        for i, st in enumerate(self.stages):
            x = st.run(x, x_ln, x_rows, feats[i], n, I_inv, E_inv, invert_intrinsic=False)
            x_rows = b
            for blk in self.layers[i]:
                x = blk.run(x)
            x_ln = None
            if i < len(self.stages) - 1:
                x, _, x_ln = self.down[i].run(x, not self.stages[i + 1].bev_flag)
        return x


class CvtDecoderPlan:
    """Decoder / DecoderBlock (nuscenes/.../model/decoder.py:6-61)."""

    def __init__(self, dec, device, dtype):
        self.blocks = []
        for blk in dec.layers:
            conv = blk.conv
            a = PackedConv(*fold_conv_bn(conv[1], conv[2]), device, dtype)
            b = PackedConv(*fold_conv_bn(conv[4], conv[5]), device, dtype)
            up = None
            if blk.up is not None:
                up = PackedConv(blk.up.weight.detach().float(), blk.up.bias.detach().float(), device, dtype)
            self.blocks.append((a, b, up))

    def run(self, x):
        skip = x
        y = x
        for a, b, up in self.blocks:
            y = ops.upsample_bilinear2x(y)
            y, _, _ = ops.conv2d_nhwc(y, a, pad=1, act=ACT_RELU)
            if up is None:
                y, _, _ = ops.conv2d_nhwc(y, b, act=ACT_RELU)
            else:
                y, _, _ = ops.conv2d_nhwc(y, b)
                u, _, _ = ops.conv2d_nhwc(skip, up)
                ops.add_nearest_relu_(y, u)
        return y


class LogitsPlan:
    """to_logits of CrossViewTransformer (nuscenes/.../model/cvt.py:29-33)."""

    def __init__(self, seq, device, dtype):
        self.a = PackedConv(*fold_conv_bn(seq[0], seq[1]), device, dtype)
        self.b = PackedConv(seq[3].weight.detach().float(), seq[3].bias.detach().float(), device, dtype)

    def run(self, y):
        y, _, _ = ops.conv2d_nhwc(y, self.a, pad=1, act=ACT_RELU)
        _, z32, _ = ops.conv2d_nhwc(y, self.b, want16=False, out32=True)
        return ops.nhwc_to_nchw(z32, torch.float32)


# --------------------------------------------------------------------------- EfficientNet-B4 prefix (MBConv)
def static_same_pad(size, k, s):
    """(before, after) padding of efficientnet_pytorch's Conv2dStaticSamePadding for the NOMINAL extent `size`."""
    out = -(-size // s)
    pad = max((out - 1) * s + k - size, 0)
    return pad // 2, pad - pad // 2


class MBConvPlan:
    """One MBConvBlock (efficientnet_pytorch 0.7.1) with eval BatchNorm (eps 1e-3) folded: 1x1 expand + swish on the
    tensor-core GEMM, depthwise kxk + swish + squeeze sums, squeeze-excite gate, gate application, 1x1 project
    (+ identity skip in the epilogue)."""

    def __init__(self, blk, k, stride, expand, cin, cout, nominal, device, dtype):
        self.k, self.stride, self.skip = k, stride, (stride == 1 and cin == cout)
        self.pad = static_same_pad(nominal, k, stride)
        self.expand = None
        if expand != 1:
            self.expand = PackedConv(*fold_conv_bn(blk._expand_conv, blk._bn0), device, dtype)
        wd, bd = fold_conv_bn(blk._depthwise_conv, blk._bn1)                      # (C, 1, k, k)
        self.wd = _dev_f32(wd[:, 0].permute(1, 2, 0), device)                      # (k, k, C)
        self.bd = _dev_f32(bd, device)
        self.w1 = _dev_f32(blk._se_reduce.weight.flatten(1), device)
        self.b1 = _dev_f32(blk._se_reduce.bias, device)
        self.w2 = _dev_f32(blk._se_expand.weight.flatten(1), device)
        self.b2 = _dev_f32(blk._se_expand.bias, device)
        self.project = PackedConv(*fold_conv_bn(blk._project_conv, blk._bn2), device, dtype)

    def run(self, x, sums_pool):
        N, H, W, _ = x.shape
        y = x
        if self.expand is not None:
            y = ops.conv2d_nhwc(y, self.expand, act=ACT_SWISH)[0]
        C = y.shape[-1]
        pb, pa = self.pad
        OH, OW = (H + pb + pa - self.k) // self.stride + 1, (W + pb + pa - self.k) // self.stride + 1
        sums = sums_pool[:N * C].view(N, C)              # zeroed once per forward for all blocks (EffNetPlan.run)
        y = ops.dwconv_swish(y, self.wd, self.bd, self.k, self.stride, pb, pb, OH, OW, sums=sums)
        ops.se_scale_(y, sums, self.w1, self.b1, self.w2, self.b2)
        return ops.conv2d_nhwc(y, self.project, res=x if self.skip else None)[0]


class EffNetPlan:
    """stem (3x3 s2 static-same conv + BN + swish on RGBx super-pixels) + the MBConv groups of
    nuscenes/cross_view_transformer/model/backbones/efficientnet.py:56-69."""

    def __init__(self, ext, device, dtype):
        self.dtype = dtype
        nominal = ext.nominal_image_size
        w, b = fold_conv_bn(ext.layers[0][0], ext.layers[0][1])                 # (48, 3, 3, 3)
        co = w.shape[0]
        # 2 pixels x 4 channels = one 16-byte super-pixel (see ResNetPlan): kernel row = 2 super-taps, stride (2, 1)
        wp = torch.zeros(co, 4, 3, 4)
        wp[:, :3, :, :3] = w
        wp = wp.reshape(co, 4, 3, 2, 2).permute(0, 4, 1, 2, 3)                   # co dx c ky kxs
        self.stem = PackedConv(wp.reshape(co, 8, 3, 2), b, device, dtype)
        self.stem.k_alg = 27
        self.stem_pad = static_same_pad(nominal, 3, 2)
        nominal = -(-nominal // 2)
        self.groups = []
        for gi, spec in enumerate(ext.block_specs):
            blocks = []
            for j, (k, s, e, cin, cout) in enumerate(spec):
                blocks.append(MBConvPlan(ext.layers[gi + 1][j], k, s, e, cin, cout, nominal, device, dtype))
                nominal = -(-nominal // s)
            self.groups.append(blocks)

    def run(self, images_nchw, mean=None, std=None):
        """images (N, 3, H, W) fp32 -> list of NHWC 16-bit group outputs."""
        N, _, H, W = images_nchw.shape
        pb, pa = self.stem_pad
        OH, OW = (H + pb + pa - 3) // 2 + 1, (W + pb + pa - 3) // 2 + 1
        Hp, Wp = max(H + pb, 2 * OH + 1), max(W + pb, 2 * OW + 2)
        Wp += Wp & 1
        x = ops.pack_rgb4_padded(images_nchw, self.dtype, pb, pb, Hp, Wp, nchw=True, mean=mean, std=std).view(N, Hp, Wp // 2, 8)
        x = ops.conv2d_nhwc(x, self.stem, stride=(2, 1), pad=0, act=ACT_SWISH, out_hw=(OH, OW))[0]
        feats = []
        # one zeroed pool serves the squeeze sums of every block (each block takes its own slice: no re-zeroing)
        need = sum(N * b.wd.shape[-1] for blocks in self.groups for b in blocks)
        pool = torch.zeros(need, dtype=torch.float32, device=x.device)
        off = 0
        for blocks in self.groups:
            for blk in blocks:
                c = N * blk.wd.shape[-1]
                x = blk.run(x, pool[off:off + c])
                off += c
            feats.append(x)
        return feats
