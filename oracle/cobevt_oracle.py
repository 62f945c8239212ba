"""CPU oracle for the CoBEVT FAX hot path -- TEST INFRASTRUCTURE ONLY.

A plain fp32 restatement (torch CPU tensor arithmetic, no nn.Modules, no einops, no
torchvision) of the reference algorithm, written from the reference's behaviour; every
function cites the reference lines it follows (paths relative to the reference root,
``opv2v/opencood/models`` abbreviated ``M``).  It is pinned against the real reference in
``tests/test_oracle_vs_reference.py`` (authoring container, where /root/reference exists)
and through the committed fixtures in ``tests/golden/`` (made by the reference itself,
``tests/golden/make_golden.py``).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this
package.  The product (cobevt_b200/) never does.

All functions take a flat ``sd`` (state dict: reference parameter names -> fp32 tensors)
plus a name prefix, so they work on any reference checkpoint.
"""
import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-5
BN_EPS = 1e-5


# --------------------------------------------------------------------------- primitives
def _lin(x, sd, name):
    w = sd[name + ".weight"]
    b = sd.get(name + ".bias")
    y = x @ w.t()
    return y if b is None else y + b


def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], LN_EPS)


def _bn(x, sd, name, eps=BN_EPS):
    """eval-mode BatchNorm2d on NCHW (running statistics)."""
    g, b = sd[name + ".weight"], sd[name + ".bias"]
    mu, var = sd[name + ".running_mean"], sd[name + ".running_var"]
    s = g / torch.sqrt(var + eps)
    return x * s[None, :, None, None] + (b - mu * s)[None, :, None, None]


def _conv(x, sd, name, stride=1, pad=0):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=pad)


# --------------------------------------------------------------------------- ResNet encoder
_RESNET_BLOCKS = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3)}


def resnet_basic_block(x, sd, p, stride):
    """torchvision BasicBlock (resnet.py:59-105): conv3x3-BN-ReLU-conv3x3-BN (+downsample) add ReLU."""
    y = F.relu(_bn(_conv(x, sd, p + ".conv1", stride, 1), sd, p + ".bn1"))
    y = _bn(_conv(y, sd, p + ".conv2", 1, 1), sd, p + ".bn2")
    if (p + ".downsample.0.weight") in sd:
        x = _bn(_conv(x, sd, p + ".downsample.0", stride, 0), sd, p + ".downsample.1")
    return F.relu(y + x)


def resnet_encoder(images, sd, prefix, num_layers, id_pick):
    """M/backbones/resnet_ms.py:46-89.  images (B, L, M, H, W, 3) -> list of (B, L, M, C, h, w)."""
    B, L, Mv, H, W, _ = images.shape
    x = images.reshape(B * L * Mv, H, W, 3).permute(0, 3, 1, 2)
    e = prefix + "encoder."
    x = F.relu(_bn(_conv(x, sd, e + "conv1", 2, 3), sd, e + "bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for li, nblk in enumerate(_RESNET_BLOCKS[num_layers]):
        for bi in range(nblk):
            x = resnet_basic_block(x, sd, f"{e}layer{li + 1}.{bi}", 2 if (li > 0 and bi == 0) else 1)
        feats.append(x.reshape(B, L, Mv, *x.shape[1:]))
    if isinstance(id_pick, (list, tuple)):
        return [feats[i] for i in id_pick]
    return feats[id_pick]


def bottleneck(x, sd, p):
    """torchvision Bottleneck(c, c//4) (resnet.py:108-160) as used at fax_modules.py:10,472."""
    y = F.relu(_bn(_conv(x, sd, p + ".conv1"), sd, p + ".bn1"))
    y = F.relu(_bn(_conv(y, sd, p + ".conv2", 1, 1), sd, p + ".bn2"))
    y = _bn(_conv(y, sd, p + ".conv3"), sd, p + ".bn3")
    return F.relu(y + x)


# --------------------------------------------------------------------------- BEV embedding
def bev_world_grids(bev_height, bev_width, h_meters, w_meters, offset, upsample_scales):
    """fax_modules.py:13-35,64-81: ego-frame (x, y, 1) of every BEV cell at each pyramid scale."""
    sh, sw = bev_height / h_meters, bev_width / w_meters
    V = torch.tensor([[0., -sw, bev_width / 2.], [-sh, 0., bev_height * offset + bev_height / 2.], [0., 0., 1.]])
    V_inv = V.inverse()
    grids = []
    for s in upsample_scales:
        h, w = bev_height // s, bev_width // s
        xs = torch.linspace(0, 1, w) * bev_width
        ys = torch.linspace(0, 1, h) * bev_height
        g = torch.stack([xs[None, :].expand(h, w), ys[:, None].expand(h, w), torch.ones(h, w)], 0)  # 3 h w
        grids.append((V_inv @ g.reshape(3, -1)).reshape(3, h, w))
    return grids


def image_plane(feat_h, feat_w, image_h, image_w):
    """fax_modules.py:13-21,275-277: pixel coordinates (u, v, 1) of every feature cell, (3, h, w)."""
    xs = torch.linspace(0, 1, feat_w) * image_w
    ys = torch.linspace(0, 1, feat_h) * image_h
    return torch.stack([xs[None, :].expand(feat_h, feat_w), ys[:, None].expand(feat_h, feat_w),
                        torch.ones(feat_h, feat_w)], 0)


# --------------------------------------------------------------------------- FAX attention
def _win_part(t, w1, w2):
    """'b n (x w1) (y w2) d -> b (x y) (n w1 w2) d' on channel-last (b, n, H, W, d)."""
    b, n, H, W, d = t.shape
    X, Y = H // w1, W // w2
    return t.reshape(b, n, X, w1, Y, w2, d).permute(0, 2, 4, 1, 3, 5, 6).reshape(b, X * Y, n * w1 * w2, d)


def _grid_part(t, w1, w2):
    """'b n (w1 x) (w2 y) d -> b (x y) (n w1 w2) d'."""
    b, n, H, W, d = t.shape
    X, Y = H // w1, W // w2
    return t.reshape(b, n, w1, X, w2, Y, d).permute(0, 3, 5, 1, 2, 4, 6).reshape(b, X * Y, n * w1 * w2, d)


def _win_merge(t, b, n, H, W, w1, w2):
    """inverse of _win_part: (b, X*Y, n*w1*w2, d) -> (b, n, H, W, d)."""
    X, Y = H // w1, W // w2
    d = t.shape[-1]
    return t.reshape(b, X, Y, n, w1, w2, d).permute(0, 3, 1, 4, 2, 5, 6).reshape(b, n, H, W, d)


def _grid_merge(t, b, n, H, W, w1, w2):
    X, Y = H // w1, W // w2
    d = t.shape[-1]
    return t.reshape(b, X, Y, n, w1, w2, d).permute(0, 3, 4, 1, 5, 2, 6).reshape(b, n, H, W, d)


def cross_win_attention(qp, kp, vp, sd, p, heads, dim_head):
    """CrossWinAttention.forward core, fax_modules.py:211-243, on partitioned tokens.

    qp: (b, windows, Lq, d)  kp, vp: (b, windows, Lk, d).  Returns proj(att) (b, windows, Lq, d)
    -- the mean over the n query views and the skip are applied by the caller (they act on
    the un-partitioned layout, :243-247).
    """
    q = _lin(_ln(qp, sd, p + ".to_q.0"), sd, p + ".to_q.1")
    k = _lin(_ln(kp, sd, p + ".to_k.0"), sd, p + ".to_k.1")
    v = _lin(_ln(vp, sd, p + ".to_v.0"), sd, p + ".to_v.1")
    b, Wn, Lq, _ = q.shape
    Lk = k.shape[2]
    sh = lambda t, L: t.reshape(b, Wn, L, heads, dim_head).permute(0, 3, 1, 2, 4)  # b m win L dh
    q, k, v = sh(q, Lq), sh(k, Lk), sh(v, Lk)
    dot = (dim_head ** -0.5) * (q @ k.transpose(-1, -2))
    a = dot.softmax(-1) @ v
    a = a.permute(0, 2, 3, 1, 4).reshape(b, Wn, Lq, heads * dim_head)
    return _lin(a, sd, p + ".proj")


def pad_divisible(x, win_h, win_w):
    """fax_modules.py:315-321 on (..., h, w): zero-pad bottom/right to a window multiple."""
    h, w = x.shape[-2:]
    ph = ((h + win_h) // win_h) * win_h - h if h % win_h else 0
    pw = ((w + win_w) // win_w) * win_w - w if w % win_w else 0
    return F.pad(x, (0, pw, 0, ph))


def cross_view_swap_attention(x, world, feature, I_inv, E_inv, sd, p, *, image_h, image_w,
                              q_win, feat_win, heads, dim_head, bev_embed_flag,
                              no_image_features=False, skip=True):
    """CrossViewSwapAttention.forward, fax_modules.py:323-441.

    x (b, d, H, W); world (2, H, W) BEV grid of this stage; feature (b, n, c, h, w);
    I_inv (b, n, 3, 3); E_inv (b, n, 4, 4).  Returns (b, d, H, W).
    """
    b, n, _, h, w = feature.shape
    d, H, W = x.shape[1:]
    dev = x.device
    pix = image_plane(h, w, image_h, image_w).reshape(3, h * w).to(dev)
    # :346-358 camera-centre and ray-direction embeddings
    c = E_inv[..., -1:].reshape(b * n, 4, 1, 1)
    c_embed = _conv(c, sd, p + ".cam_embed")                                   # (bn, d, 1, 1)
    cam = I_inv @ pix                                                           # b n 3 hw
    cam = torch.cat([cam, torch.ones(b, n, 1, h * w, device=dev)], 2)
    dirs = (E_inv @ cam).reshape(b * n, 4, h, w)
    img_embed = _conv(dirs, sd, p + ".img_embed") - c_embed
    img_embed = img_embed / (img_embed.norm(dim=1, keepdim=True) + 1e-7)
    # :370-375 BEV positional query offsets (first stage only)
    if bev_embed_flag:
        w_embed = _conv(world[None], sd, p + ".bev_embed")                      # 1 d H W
        bev_embed = w_embed - c_embed
        bev_embed = bev_embed / (bev_embed.norm(dim=1, keepdim=True) + 1e-7)
        query = bev_embed.reshape(b, n, d, H, W) + x[:, None]
    else:
        query = x[:, None]
    # :377-384 key / value maps
    feat = feature.reshape(b * n, *feature.shape[2:])
    proj = lambda name: _conv(F.relu(_bn(feat, sd, p + f".{name}.0")), sd, p + f".{name}.2")
    key = img_embed if no_image_features else img_embed + proj("feature_proj")
    val = proj("feature_linear")
    key = pad_divisible(key.reshape(b, n, d, h, w), *feat_win)                  # :395-396
    val = pad_divisible(val.reshape(b, n, d, h, w), *feat_win)
    cl = lambda t: t.permute(0, 1, 3, 4, 2)                                     # channel last
    q_cl, k_cl, v_cl = cl(query), cl(key), cl(val)
    nq = q_cl.shape[1]
    x_cl = x.permute(0, 2, 3, 1)
    # :399-409 local window <-> local window
    z = cross_win_attention(_win_part(q_cl, *q_win), _win_part(k_cl, *feat_win), _win_part(v_cl, *feat_win),
                            sd, p + ".cross_win_attend_1", heads, dim_head)
    z = _win_merge(z, b, nq, H, W, *q_win).mean(1)
    if skip:
        z = z + x_cl
    z = z + _lin(F.gelu(_lin(_ln(z, sd, p + ".prenorm_1"), sd, p + ".mlp_1.0")), sd, p + ".mlp_1.2")  # :411
    # :413-433 local window <-> dilated grid.  The reference repeats the query n times and
    # averages n identical results (:414, :243); one copy is the same function.
    z2 = cross_win_attention(_win_part(z[:, None], *q_win), _grid_part(k_cl, *feat_win), _grid_part(v_cl, *feat_win),
                             sd, p + ".cross_win_attend_2", heads, dim_head)
    z2 = _win_merge(z2, b, 1, H, W, *q_win)[:, 0]
    if skip:
        z2 = z2 + z
    z2 = z2 + _lin(F.gelu(_lin(_ln(z2, sd, p + ".prenorm_2"), sd, p + ".mlp_2.0")), sd, p + ".mlp_2.2")  # :435
    z2 = _ln(z2, sd, p + ".postnorm")
    return z2.permute(0, 3, 1, 2)


def rel_pos_index_2d(ws):
    """fax_modules.py:123-128."""
    pos = torch.arange(ws)
    yy, xx = torch.meshgrid(pos, pos, indexing="ij")
    g = torch.stack([yy.reshape(-1), xx.reshape(-1)], 1)
    rel = g[:, None, :] - g[None, :, :] + ws - 1
    return rel[..., 0] * (2 * ws - 1) + rel[..., 1]


def global_attention(x, sd, p, dim_head, window_size):
    """fax_modules.Attention.forward :132-176 (eval: dropout is identity)."""
    b, d, H, W = x.shape
    heads = d // dim_head
    t = x.permute(0, 2, 3, 1).reshape(b, H * W, d)
    q, k, v = _lin(t, sd, p + ".to_qkv").chunk(3, -1)
    sh = lambda u: u.reshape(b, H * W, heads, dim_head).permute(0, 2, 1, 3)
    q, k, v = sh(q) * dim_head ** -0.5, sh(k), sh(v)
    bias = sd[p + ".rel_pos_bias.weight"][rel_pos_index_2d(window_size).to(x.device)]        # T T heads
    sim = q @ k.transpose(-1, -2) + bias.permute(2, 0, 1)
    out = sim.softmax(-1) @ v
    out = out.permute(0, 2, 1, 3).reshape(b, H, W, d)
    return _lin(out, sd, p + ".to_out.0").permute(0, 3, 1, 2)


def fax_module(features, intrinsic, extrinsic, sd, prefix, cfg):
    """FAXModule.forward, fax_modules.py:497-521.

    features: list of (b, l, n, c, h, w); intrinsic (b, l, n, 3, 3); extrinsic (b, l, n, 4, 4).
    Returns (b, l, d, H', W').
    """
    b, l, n = features[0].shape[:3]
    I_inv = intrinsic.reshape(b * l, n, 3, 3).inverse()
    E_inv = extrinsic.reshape(b * l, n, 4, 4)           # used un-inverted (:502-503)
    be = cfg["bev_embedding"]
    grids = bev_world_grids(be["bev_height"], be["bev_width"], be["h_meters"], be["w_meters"], be["offset"],
                            be["upsample_scales"])
    x = sd[prefix + "bev_embedding.learned_features"][None].expand(b * l, -1, -1, -1)
    cv, sw = cfg["cross_view"], cfg["cross_view_swap"]
    for i, feat in enumerate(features):
        f = feat.reshape(b * l, *feat.shape[2:])
        x = cross_view_swap_attention(
            x, grids[i][:2].to(x.device), f, I_inv, E_inv, sd, f"{prefix}cross_views.{i}",
            image_h=cv["image_height"], image_w=cv["image_width"], q_win=sw["q_win_size"][i],
            feat_win=sw["feat_win_size"][i], heads=cv["heads"][i], dim_head=cv["dim_head"][i],
            bev_embed_flag=sw["bev_embedding_flag"][i], no_image_features=cv["no_image_features"],
            skip=cv["skip"])
        for j in range(cfg["middle"][i]):
            x = bottleneck(x, sd, f"{prefix}layers.{i}.{j}")
        if i < len(features) - 1:                        # fax_modules.py:476-489
            dp = f"{prefix}downsample_layers.{i}.0"
            x = F.pixel_unshuffle(_conv(x, sd, dp + ".0", 1, 1), 2)
            x = F.relu(_bn(_conv(x, sd, dp + ".2", 1, 1), sd, dp + ".3"))
            x = _bn(_conv(x, sd, dp + ".5"), sd, dp + ".6")
    sa = cfg["self_attn"]
    x = global_attention(x, sd, prefix + "self_attn", sa["dim_head"], sa["window_size"])
    return x.reshape(b, l, *x.shape[1:])


# --------------------------------------------------------------------------- FuseBEVT
def rel_pos_index_3d(L, wh, ww):
    """swap_fusion_modules.py:63-83 (torch.meshgrid default indexing = 'ij')."""
    cd, ch, cw = torch.meshgrid(torch.arange(L), torch.arange(wh), torch.arange(ww), indexing="ij")
    c = torch.stack([cd.reshape(-1), ch.reshape(-1), cw.reshape(-1)], 0)
    rel = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0).clone()
    rel[..., 0] += L - 1
    rel[..., 1] += wh - 1
    rel[..., 2] += ww - 1
    return rel[..., 0] * (2 * wh - 1) * (2 * ww - 1) + rel[..., 1] * (2 * ww - 1) + rel[..., 2]


def swap_attention(xp, mask_p, sd, p, dim_head, L, ws):
    """swap_fusion_modules.Attention.forward :87-128 on partitioned tokens.

    xp: (B*, T, d) with T = L*ws*ws ordered (l, w1, w2); mask_p: (B*, T) or None (1 = keep).
    """
    Bw, T, d = xp.shape
    heads = d // dim_head
    q, k, v = _lin(xp, sd, p + ".to_qkv").chunk(3, -1)
    sh = lambda u: u.reshape(Bw, T, heads, dim_head).permute(0, 2, 1, 3)
    q, k, v = sh(q) * dim_head ** -0.5, sh(k), sh(v)
    bias = sd[p + ".relative_position_bias_table.weight"][rel_pos_index_3d(L, ws, ws).to(xp.device)]
    sim = q @ k.transpose(-1, -2) + bias.permute(2, 0, 1)
    if mask_p is not None:
        sim = sim.masked_fill(mask_p[:, None, None, :] == 0, float("-inf"))
    out = sim.softmax(-1) @ v
    out = out.permute(0, 2, 1, 3).reshape(Bw, T, d)
    return _lin(out, sd, p + ".to_out.0")


def swap_fusion_block(x, mask, sd, p, dim_head, ws):
    """SwapFusionBlockMask.forward :165-192 / SwapFusionBlock :209-230 (mask=None).

    x (B, L, d, H, W); mask (B, H, W, 1, L) or None.
    """
    B, L, d, H, W = x.shape
    t = x.permute(0, 1, 3, 4, 2)                                                # B L H W d
    for kind, part, merge in (("window", _win_part, _win_merge), ("grid", _grid_part, _grid_merge)):
        xp = part(t, ws, ws)                                                    # B win T d
        Wn, T = xp.shape[1:3]
        mp = None
        if mask is not None:
            mm = mask[:, :, :, 0, :].permute(0, 3, 1, 2)[..., None]             # B L H W 1
            mp = part(mm, ws, ws).reshape(B * Wn, T)
        a = f"{p}.{kind}_attention"
        y = swap_attention(_ln(xp, sd, a + ".norm").reshape(B * Wn, T, d), mp, sd, a + ".fn", dim_head, L, ws)
        xp = xp + y.reshape(B, Wn, T, d)
        f = f"{p}.{kind}_ffd"
        xp = xp + _lin(F.gelu(_lin(_ln(xp, sd, f + ".norm"), sd, f + ".fn.net.0")), sd, f + ".fn.net.3")
        t = merge(xp, B, L, H, W, ws, ws)
    return t.permute(0, 1, 4, 2, 3)


def swap_fusion_encoder(x, mask, sd, prefix, args):
    """SwapFusionEncoder.forward :283-286 with the mlp_head :275-281."""
    use_mask = args.get("mask", False)
    for i in range(args["depth"]):
        if use_mask:
            x = swap_fusion_block(x, mask, sd, f"{prefix}layers.{i}", args["dim_head"], args["window_size"])
        else:   # SwapFusionBlock keeps its sub-modules in an nn.Sequential named 'block'
            x = _swap_fusion_block_seq(x, sd, f"{prefix}layers.{i}.block", args["dim_head"], args["window_size"])
    t = x.mean(1).permute(0, 2, 3, 1)                                           # mean over ALL agent slots
    t = _lin(_ln(t, sd, prefix + "mlp_head.2"), sd, prefix + "mlp_head.3")
    return t.permute(0, 3, 1, 2)


def _swap_fusion_block_seq(x, sd, p, dim_head, ws):
    """SwapFusionBlock (:209-225): Sequential indices 1,2 (window attn, ffd) and 5,6 (grid)."""
    B, L, d, H, W = x.shape
    t = x.permute(0, 1, 3, 4, 2)
    for (ia, iff), part, merge in (((1, 2), _win_part, _win_merge), ((5, 6), _grid_part, _grid_merge)):
        xp = part(t, ws, ws)
        Wn, T = xp.shape[1:3]
        y = swap_attention(_ln(xp, sd, f"{p}.{ia}.norm").reshape(B * Wn, T, d), None, sd, f"{p}.{ia}.fn", dim_head, L, ws)
        xp = xp + y.reshape(B, Wn, T, d)
        xp = xp + _lin(F.gelu(_lin(_ln(xp, sd, f"{p}.{iff}.norm"), sd, f"{p}.{iff}.fn.net.0")), sd, f"{p}.{iff}.fn.net.3")
        t = merge(xp, B, L, H, W, ws, ws)
    return t.permute(0, 1, 4, 2, 3)


# --------------------------------------------------------------------------- regroup / STTF
def regroup(x, record_len, max_len):
    """M/sub_modules/fuse_utils.py:8-61: (sum L, C, H, W) -> (B, max_len, C, H, W), mask (B, max_len)."""
    out, mask, s = [], [], 0
    for n in [int(v) for v in record_len]:
        f = x[s:s + n]
        s += n
        out.append(torch.cat([f, torch.zeros(max_len - n, *x.shape[1:], device=x.device)], 0))
        mask.append([1] * n + [0] * (max_len - n))
    return torch.stack(out, 0), torch.tensor(mask, device=x.device)


def _sttf_pixel_matrix(tm, resolution, downsample_rate, H, W):
    """The 3x3 pixel-space matrix T (dst <- src) the reference hands to warp_affine:
    torch_transformation_utils.py:108-134 (discretise) and :254-297 (rotate about the centre)."""
    m = tm[..., [0, 1], :][..., [0, 1, 3]].clone().float()                       # (..., 2, 3)
    dev = tm.device
    m[..., 2] = m[..., 2] / (resolution * downsample_rate)
    m = m.reshape(-1, 2, 3)
    N = m.shape[0]
    cx, cy = W / 2, H / 2
    sh = torch.eye(3, device=dev).repeat(N, 1, 1)
    sh[:, 0, 2], sh[:, 1, 2] = cx, cy
    shi = torch.eye(3, device=dev).repeat(N, 1, 1)
    shi[:, 0, 2], shi[:, 1, 2] = -cx, -cy
    rot = torch.eye(3, device=dev).repeat(N, 1, 1)
    rot[:, :2, :2] = m[:, :, :2]
    T = sh @ rot @ shi
    T[:, :2, 2] += m[:, :, 2]
    return T


def warp_affine_pixels(src, T, mode):
    """warp_affine (torch_transformation_utils.py:317-355) restated in pixel coordinates:
    with align_corners=True the normalise / affine_grid / grid_sample chain samples
    src at T^-1 (x_dst, y_dst, 1), zeros outside.  src (N, C, H, W), T (N, 3, 3)."""
    N, C, H, W = src.shape
    Ti = T.float().inverse().to(T.dtype)   # (fp32 solve: linalg.inv has no half kernel; only the eager-autocast proxy ever gets here in 16 bit)
    dev = src.device
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev), torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
    p = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, device=dev)], 0)       # 3, HW
    s = Ti @ p                                                                    # N 3 HW
    sx, sy = s[:, 0], s[:, 1]
    flat = src.reshape(N, C, H * W)

    def tap(ix, iy):
        ok = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
        idx = (iy.clamp(0, H - 1) * W + ix.clamp(0, W - 1)).long()
        return flat.gather(2, idx[:, None, :].expand(N, C, H * W)) * ok[:, None, :].float()

    if mode == "nearest":
        out = tap(torch.round(sx), torch.round(sy))     # round-half-even, like grid_sample
    else:
        x0, y0 = torch.floor(sx), torch.floor(sy)
        fx, fy = (sx - x0)[:, None], (sy - y0)[:, None]
        out = (tap(x0, y0) * (1 - fx) * (1 - fy) + tap(x0 + 1, y0) * fx * (1 - fy)
               + tap(x0, y0 + 1) * (1 - fx) * fy + tap(x0 + 1, y0 + 1) * fx * fy)
    return out.reshape(N, C, H, W)


def sttf(x, tm, resolution, downsample_rate):
    """STTF.forward, M/corpbevt.py:28-64: (B, L, C, H, W) -> (B, L, H, W, C) in the ego frame."""
    B, L, C, H, W = x.shape
    t = x.transpose(3, 4).flip(4)                                                # 'b l c w h', flipped
    Hh, Ww = t.shape[3:]
    T = _sttf_pixel_matrix(tm, resolution, downsample_rate, Hh, Ww)
    y = warp_affine_pixels(t.reshape(B * L, C, Hh, Ww), T, "bilinear").reshape(B, L, C, Hh, Ww)
    return y.flip(4).permute(0, 1, 4, 3, 2)                                      # 'b l c w h -> b l h w c'


def roi_and_cav_mask(shape, cav_mask, tm, resolution, downsample_rate):
    """get_roi_and_cav_mask, torch_transformation_utils.py:11-49 -> (B, H, W, 1, L)."""
    B, L, H, W, _ = shape
    T = _sttf_pixel_matrix(tm, resolution, downsample_rate, H, W)
    roi = warp_affine_pixels(torch.ones(B * L, 1, H, W, device=tm.device), T, "nearest").reshape(B, L, 1, H, W)
    com = roi * cav_mask[:, :, None, None, None].float()
    return com.permute(0, 3, 4, 2, 1)


# --------------------------------------------------------------------------- decoder / head
def naive_decoder(x, sd, prefix, num_layer):
    """M/sub_modules/naive_decoder.py:62-91; parameters live in decoder.{0..6*num_layer-1}."""
    b, l = x.shape[:2]
    t = x.reshape(b * l, *x.shape[2:])
    for j in range(num_layer):
        k = 6 * j
        t = F.relu(_bn(_conv(t, sd, f"{prefix}decoder.{k}", 1, 1), sd, f"{prefix}decoder.{k + 1}"))
        t = F.interpolate(t, scale_factor=2, mode="nearest")
        t = F.relu(_bn(_conv(t, sd, f"{prefix}decoder.{k + 3}", 1, 1), sd, f"{prefix}decoder.{k + 4}"))
    return t.reshape(b, l, *t.shape[1:])


def bev_seg_head(x, sd, prefix, target, b, l):
    """M/sub_modules/bev_seg_head.py:35-61."""
    out = {}
    if target != "static":         # 'dynamic' and every other target string run the dynamic head (:36-39,50-53)
        out["dynamic_seg"] = _conv(x, sd, prefix + "dynamic_head", 1, 1).reshape(b, l, -1, *x.shape[2:])
    if target != "dynamic":
        out["static_seg"] = _conv(x, sd, prefix + "static_head", 1, 1).reshape(b, l, -1, *x.shape[2:])
    if target == "dynamic":
        out["static_seg"] = torch.zeros_like(out["dynamic_seg"])
    if target == "static":
        out["dynamic_seg"] = torch.zeros_like(out["static_seg"])
    return out


def naive_compressor(x, sd, prefix):
    """NaiveCompressor.forward, M/sub_modules/naive_compress.py:5-28: conv3x3-BN-ReLU down to C/ratio, then two
    conv3x3-BN-ReLU back up; every BatchNorm has eps = 1e-3 (:11,17,20)."""
    y = F.relu(_bn(_conv(x, sd, prefix + "encoder.0", 1, 1), sd, prefix + "encoder.1", eps=1e-3))
    y = F.relu(_bn(_conv(y, sd, prefix + "decoder.0", 1, 1), sd, prefix + "decoder.1", eps=1e-3))
    return F.relu(_bn(_conv(y, sd, prefix + "decoder.3", 1, 1), sd, prefix + "decoder.4", eps=1e-3))


def post_process(output_dict):
    """CameraBevPostprocessor.post_process_train, opv2v/opencood/data_utils/post_processor/
    camera_bev_postprocessor.py:55-94: per head softmax over the class axis of seg[:, 0] and its argmax."""
    out = dict(output_dict)
    for key in ("static", "dynamic"):
        seg = output_dict[key + "_seg"][:, 0]
        prob = torch.softmax(seg, dim=1)
        out[key + "_prob"] = prob
        out[key + "_map"] = torch.argmax(prob, dim=1)
    return out


# --------------------------------------------------------------------------- whole models
def corpbevt(batch, sd, cfg):
    """CorpBEVT.forward, M/corpbevt.py:104-145."""
    x = batch["inputs"]
    enc = cfg["encoder"]
    feats = resnet_encoder(x, sd, "encoder.", enc["num_layers"], enc["id_pick"])
    y = fax_module(feats, batch["intrinsic"], batch["extrinsic"], sd, "fax.", cfg["fax"])
    y = y.squeeze(1)
    if cfg.get("compression", 0) > 0:                                            # :118-121
        y = naive_compressor(y, sd, "naive_compressor.")
    y, mask = regroup(y, batch["record_len"].tolist(), cfg["max_cav"])
    st = cfg["sttf"]
    tm = batch["transformation_matrix"]
    y = sttf(y, tm, st["resolution"], st["downsample_rate"])                     # B L H W C
    if st["use_roi_mask"]:
        com_mask = roi_and_cav_mask(y.shape, mask, tm, st["resolution"], st["downsample_rate"])
    else:
        com_mask = mask[:, None, None, None, :]
    y = swap_fusion_encoder(y.permute(0, 1, 4, 2, 3), com_mask, sd, "fusion_net.", cfg["fax_fusion"])
    y = naive_decoder(y[:, None], sd, "decoder.", cfg["decoder"]["num_layer"])
    y = y.reshape(-1, *y.shape[2:])
    return bev_seg_head(y, sd, "seg_head.", cfg["target"], y.shape[0], 1)


def fax_fused_transformer(batch, sd, cfg):
    """FaxFusedTransformer.forward, M/fax_fused_transformer.py:34-48."""
    x = batch["inputs"]
    b, l = x.shape[:2]
    enc = cfg["encoder"]
    feats = resnet_encoder(x, sd, "encoder.", enc["num_layers"], enc["id_pick"])
    y = fax_module(feats, batch["intrinsic"], batch["extrinsic"], sd, "fax.", cfg["fax"])
    y = naive_decoder(y, sd, "decoder.", cfg["decoder"]["num_layer"])
    y = y.reshape(b * l, *y.shape[2:])
    return bev_seg_head(y, sd, "seg_head.", cfg["target"], b, l)


# --------------------------------------------------------------------------- nuScenes SinBEVT (CVT fork)
def pyramid_axial_encoder(features, intrinsics, extrinsics, sd, prefix, cfg):
    """PyramidAxialEncoder.forward after the backbone, nuscenes/.../encoder_pyramid_axial.py:534-558.

    features: list of (b*n, c, h, w); intrinsics (b, n, 3, 3), extrinsics (b, n, 4, 4) -- BOTH inverted
    here (:538-539).  cfg: the constructor kwargs (cross_view, cross_view_swap, bev_embedding, dim, middle).
    """
    b, n = intrinsics.shape[:2]
    I_inv, E_inv = intrinsics.inverse(), extrinsics.inverse()
    be = cfg["bev_embedding"]
    grids = bev_world_grids(be["bev_height"], be["bev_width"], be["h_meters"], be["w_meters"], be["offset"],
                            be["upsample_scales"])
    x = sd[prefix + "bev_embedding.learned_features"][None].expand(b, -1, -1, -1)
    cv, sw = cfg["cross_view"], cfg["cross_view_swap"]
    for i, feat in enumerate(features):
        f = feat.reshape(b, n, *feat.shape[1:])
        x = cross_view_swap_attention(
            x, grids[i][:2].to(x.device), f, I_inv, E_inv, sd, f"{prefix}cross_views.{i}",
            image_h=cv["image_height"], image_w=cv["image_width"], q_win=sw["q_win_size"][i],
            feat_win=sw["feat_win_size"][i], heads=cv["heads"][i], dim_head=cv["dim_head"][i],
            bev_embed_flag=sw["bev_embedding_flag"][i], no_image_features=cv["no_image_features"], skip=cv["skip"])
        for j in range(cfg["middle"][i]):
            x = bottleneck(x, sd, f"{prefix}layers.{i}.{j}")
        if i < len(features) - 1:                        # encoder_pyramid_axial.py:513-527 (dim -> dim/2 -> unshuffle)
            dp = f"{prefix}downsample_layers.{i}.0"
            x = F.pixel_unshuffle(_conv(x, sd, dp + ".0", 1, 1), 2)
            x = F.relu(_bn(_conv(x, sd, dp + ".2", 1, 1), sd, dp + ".3"))
            x = _bn(_conv(x, sd, dp + ".5"), sd, dp + ".6")
    return x


def cvt_decoder(x, sd, prefix, n_blocks):
    """Decoder.forward, nuscenes/.../decoder.py:55-61 with DecoderBlock.forward :28-36."""
    y = x
    for i in range(n_blocks):
        p = f"{prefix}layers.{i}"
        t = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=True)
        t = F.relu(_bn(_conv(t, sd, p + ".conv.1", 1, 1), sd, p + ".conv.2"))
        t = _bn(_conv(t, sd, p + ".conv.4"), sd, p + ".conv.5")
        if (p + ".up.weight") in sd:
            up = F.interpolate(_conv(x, sd, p + ".up"), t.shape[-2:])          # nearest
            t = t + up
        y = F.relu(t)
    return y


def resnet_levels(image_nchw, sd, prefix, num_layers=34, n_levels=3):
    """torchvision ResNet levels 1..n on an NCHW image (the backbone adaptor of SURVEY 8d (ii))."""
    x = F.relu(_bn(_conv(image_nchw, sd, prefix + "conv1", 2, 3), sd, prefix + "bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for li, nblk in enumerate(_RESNET_BLOCKS[num_layers][:n_levels]):
        for bi in range(nblk):
            x = resnet_basic_block(x, sd, f"{prefix}layer{li + 1}.{bi}", 2 if (li > 0 and bi == 0) else 1)
        feats.append(x)
    return feats


def cross_view_transformer(batch, sd, cfg, backbone_fn):
    """CrossViewTransformer.forward, nuscenes/.../cvt.py:35-40.  backbone_fn(normalised (b*n,3,H,W)) -> features."""
    image = batch["image"]
    b, n = image.shape[:2]
    mean = torch.tensor([0.485, 0.456, 0.406], device=image.device)[None, :, None, None]
    std = torch.tensor([0.229, 0.224, 0.225], device=image.device)[None, :, None, None]
    feats = backbone_fn((image.flatten(0, 1) - mean) / std)
    x = pyramid_axial_encoder(feats, batch["intrinsics"], batch["extrinsics"], sd, "encoder.", cfg["encoder"])
    y = cvt_decoder(x, sd, "decoder.", len(cfg["decoder_blocks"]))
    z = F.relu(_bn(_conv(y, sd, "to_logits.0", 1, 1), sd, "to_logits.1"))
    z = _conv(z, sd, "to_logits.3")
    return {k: z[:, a:b_] for k, (a, b_) in cfg["outputs"].items()}


# --------------------------------------------------------------------------- EfficientNet-B4 prefix (nuScenes backbone)
# PARITY UNPINNED: the arithmetic lives in the third-party package efficientnet_pytorch==0.7.1
# (nuscenes/requirements.txt:13), which is neither vendored under /root/reference nor installed.  What follows
# restates that package's published MBConvBlock / Conv2dStaticSamePadding semantics for exactly the layers the
# reference runs (nuscenes/cross_view_transformer/model/backbones/efficientnet.py:56-69: stem + blocks 0..10), and
# is pinned STRUCTURALLY by strict-loading the 236 backbone tensors of the shipped checkpoint
# (tests/test_effnet_cpu.py); its numbers cannot be checked against the package here.

# (kernel, stride, expand_ratio, in, out) of MBConv blocks 0..10 of EfficientNet-B4 (width 1.4, depth 1.8,
# divisor 8), grouped as efficientnet.py:15-20 slices them: (0,3), (3,7), (7,11)
EFFNET_B4_PREFIX = (
    ((3, 1, 1, 48, 24), (3, 1, 1, 24, 24), (3, 2, 6, 24, 32)),
    ((3, 1, 6, 32, 32), (3, 1, 6, 32, 32), (3, 1, 6, 32, 32), (5, 2, 6, 32, 56)),
    ((5, 1, 6, 56, 56), (5, 1, 6, 56, 56), (5, 1, 6, 56, 56), (3, 2, 6, 56, 112)),
)
EFFNET_BN_EPS = 1e-3
EFFNET_B4_IMAGE_SIZE = 380        # Conv2dStaticSamePadding pads for the model's NOMINAL resolution, not the input's


def static_same_pad(size, k, s):
    """(before, after) zero padding of efficientnet_pytorch's Conv2dStaticSamePadding for nominal extent `size`."""
    out = -(-size // s)
    pad = max((out - 1) * s + k - size, 0)
    return pad // 2, pad - pad // 2


def _swish(x):
    return x * torch.sigmoid(x)


def _conv_static_same(x, w, stride, nominal):
    k = w.shape[-1]
    b, a = static_same_pad(nominal, k, stride)
    return F.conv2d(F.pad(x, (b, a, b, a)), w, None, stride=stride, groups=w.shape[0] if w.shape[1] == 1 else 1)


def mbconv_block(x, sd, p, k, stride, expand, cin, cout, nominal):
    """MBConvBlock.forward (eval: no drop-connect): expand 1x1-BN-swish, depthwise kxk-BN-swish, squeeze-excite
    (avg-pool -> 1x1 -> swish -> 1x1 -> sigmoid gate), project 1x1-BN, identity skip when stride 1 and cin == cout."""
    y = x
    if expand != 1:
        y = _swish(_bn(F.conv2d(y, sd[p + "._expand_conv.weight"]), sd, p + "._bn0", EFFNET_BN_EPS))
    y = _swish(_bn(_conv_static_same(y, sd[p + "._depthwise_conv.weight"], stride, nominal), sd, p + "._bn1", EFFNET_BN_EPS))
    sq = y.mean((2, 3), keepdim=True)
    sq = F.conv2d(_swish(F.conv2d(sq, sd[p + "._se_reduce.weight"], sd[p + "._se_reduce.bias"])),
                  sd[p + "._se_expand.weight"], sd[p + "._se_expand.bias"])
    y = torch.sigmoid(sq) * y
    y = _bn(F.conv2d(y, sd[p + "._project_conv.weight"]), sd, p + "._bn2", EFFNET_BN_EPS)
    if stride == 1 and cin == cout:
        y = y + x
    return y


def efficientnet_b4_prefix(image_nchw, sd, prefix):
    """EfficientNetExtractor.forward (efficientnet.py:80-94) for layer_names reduction_2..4: stem + MBConv groups
    (0,3), (3,7), (7,11); returns the three group outputs (32 / 56 / 112 channels at strides 4 / 8 / 16).
    Parameters: `prefix`layers.0.{0,1} (stem conv, BN), `prefix`layers.{1,2,3}.{j}._* (blocks)."""
    nominal = EFFNET_B4_IMAGE_SIZE
    x = _swish(_bn(_conv_static_same(image_nchw, sd[prefix + "layers.0.0.weight"], 2, nominal), sd, prefix + "layers.0.1",
                   EFFNET_BN_EPS))
    nominal = -(-nominal // 2)
    feats = []
    for gi, group in enumerate(EFFNET_B4_PREFIX):
        for j, (k, s, e, cin, cout) in enumerate(group):
            x = mbconv_block(x, sd, f"{prefix}layers.{gi + 1}.{j}", k, s, e, cin, cout, nominal)
            nominal = -(-nominal // s)
        feats.append(x)
    return feats
