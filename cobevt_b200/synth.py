"""Deterministic synthetic weights and inputs: no dataset or checkpoint is needed,
and the same function regenerates identical tensors on any machine with this torch build, so
golden outputs made from the reference can be checked on the GPU box without the reference.
"""
import math
import zlib

import torch


def _gen(seed, name):
    g = torch.Generator()
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
    return g


def synth_state_dict(module, seed=0):
    """Values for every floating-point entry of module.state_dict() (integer buffers keep the
    module's own value).  BN running stats / affines and LayerNorm affines are perturbed away
    from their identity defaults so folding bugs cannot hide."""
    out = {}
    for name, t in module.state_dict().items():
        if not t.is_floating_point():
            out[name] = t.clone()
            continue
        g = _gen(seed, name)
        shp = tuple(t.shape)
        if name.endswith("running_var"):
            v = torch.rand(shp, generator=g) + 0.5
        elif name.endswith("running_mean"):
            v = torch.randn(shp, generator=g) * 0.1
        elif t.dim() == 1 and name.endswith(".weight"):          # norm scales
            v = torch.rand(shp, generator=g) + 0.5
        elif t.dim() == 1:                                        # biases
            v = torch.randn(shp, generator=g) * 0.1
        elif name.endswith("learned_features"):
            v = torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            v = torch.randn(shp, generator=g) * (1.0 / math.sqrt(max(fan_in, 1)))
            if "bias_table" in name or "rel_pos_bias" in name:
                v = torch.randn(shp, generator=g) * 0.5
        out[name] = v.to(t.dtype)
    return out


def _rot_z(yaw):
    c, s = math.cos(yaw), math.sin(yaw)
    return torch.tensor([[c, -s, 0.], [s, c, 0.], [0., 0., 1.]])


def synth_batch(cfg, record_len=(5,), n_cams=4, seed=1, fov_deg=100.0):
    """A CorpBEVT batch dict: inputs (sumL,1,M,H,W,3) ~ N(0,1); pinhole intrinsics with 1% noise;
    rigid cam->ego extrinsics (yaw 360*k/M); agent->ego SE(2) transforms, identity for ego,
    identity-padded to max_cav."""
    g = torch.Generator().manual_seed(seed)
    enc = cfg['encoder']
    H, W = enc['image_height'], enc['image_width']
    L = int(sum(record_len))
    B = len(record_len)
    inputs = torch.randn(L, 1, n_cams, H, W, 3, generator=g)
    f = W / (2 * math.tan(math.radians(fov_deg) / 2))
    K = torch.tensor([[f, 0., W / 2], [0., f, H / 2], [0., 0., 1.]])
    intrinsic = K.expand(L, 1, n_cams, 3, 3).clone()
    intrinsic = intrinsic * (1 + 0.01 * torch.randn(L, 1, n_cams, 3, 3, generator=g))
    intrinsic[..., 2, :] = torch.tensor([0., 0., 1.])
    extrinsic = torch.zeros(L, 1, n_cams, 4, 4)
    # camera axes (x right, y down, z forward) expressed in the ego frame (x forward, y left, z up)
    cam2ego = torch.tensor([[0., 0., 1.], [-1., 0., 0.], [0., -1., 0.]])
    for a in range(L):
        for k in range(n_cams):
            R = _rot_z(2 * math.pi * k / n_cams) @ cam2ego
            t = (torch.rand(3, generator=g) - 0.5) * 3.0
            extrinsic[a, 0, k, :3, :3] = R
            extrinsic[a, 0, k, :3, 3] = t
            extrinsic[a, 0, k, 3, 3] = 1.0
    max_cav = cfg['max_cav']
    tm = torch.eye(4).repeat(B, max_cav, 1, 1)
    for b, n in enumerate(record_len):
        for l in range(1, n):
            yaw = (torch.rand(1, generator=g).item() * 2 - 1) * math.pi
            tm[b, l, :3, :3] = _rot_z(yaw)
            tm[b, l, :2, 3] = (torch.rand(2, generator=g) * 2 - 1) * 20.0
    return {'inputs': inputs, 'intrinsic': intrinsic, 'extrinsic': extrinsic,
            'transformation_matrix': tm, 'record_len': torch.tensor(list(record_len), dtype=torch.int64)}


def to_device(batch, device):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in batch.items()}


def synth_nuscenes_batch(image_h, image_w, b=1, n_cams=6, seed=1, fov_deg=70.0):
    """nuScenes-style batch: image (b,n,3,H,W) in [0,1); intrinsics (b,n,3,3); extrinsics (b,n,4,4) =
    ego->camera rigid transforms (the model inverts them, encoder_pyramid_axial.py:539)."""
    g = torch.Generator().manual_seed(seed)
    image = torch.rand(b, n_cams, 3, image_h, image_w, generator=g)
    f = image_w / (2 * math.tan(math.radians(fov_deg) / 2))
    K = torch.tensor([[f, 0., image_w / 2], [0., f, image_h / 2], [0., 0., 1.]])
    intr = K.expand(b, n_cams, 3, 3).clone() * (1 + 0.01 * torch.randn(b, n_cams, 3, 3, generator=g))
    intr[..., 2, :] = torch.tensor([0., 0., 1.])
    cam2ego = torch.tensor([[0., 0., 1.], [-1., 0., 0.], [0., -1., 0.]])
    extr = torch.zeros(b, n_cams, 4, 4)
    for i in range(b):
        for k in range(n_cams):
            T = torch.eye(4)
            T[:3, :3] = _rot_z(2 * math.pi * k / n_cams) @ cam2ego
            T[:3, 3] = (torch.rand(3, generator=g) - 0.5) * 3.0
            extr[i, k] = torch.linalg.inv(T)
    return {'image': image, 'intrinsics': intr, 'extrinsics': extr}
