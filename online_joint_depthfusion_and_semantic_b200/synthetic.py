"""Seeded synthetic RGB-D scenes for the fusion hot path.

No dataset and no network exist in this environment, so every test and bench
feeds the path frames rendered from an analytic signed-distance scene:
a box room with a sphere and a cuboid in it.  The generator emits exactly the
`sample` dict the reference loaders produce (dataset/replica.py:211-295):
`image` (3,h,w) f32 normalised with the Replica mean/std, `<input>` depth
(h,w) f32 in metres, `mask` = 0.05 < z < 5 (dataset/replica.py:255),
`extrinsics` cam->world (4,4) f32 with z forward / y down / x right
(dataset/replica.py:268-279), `intrinsics` (3,3) f64 with f = cx = cy = h/2
for the Replica convention (dataset/replica.py:281-288), `semantic_gt` u8,
`frame_id` "scene/traj/idx".

It is host-side input plumbing (torch ops on whatever device is asked for),
not part of the measured path.
"""
import math

import numpy as np
import torch

ROOM_HALF = 1.4
SPHERE_C = (0.35, -0.25, -0.55)
SPHERE_R = 0.5
BOX_C = (-0.55, 0.45, -0.95)
BOX_H = (0.35, 0.3, 0.45)

# label ids (<= 29 so they fit the 30-class Replica head)
LABEL_WALLS = (1, 2, 3, 4, 5, 6)
LABEL_SPHERE = 7
LABEL_BOX = 8

_PALETTE = np.array(
    [[0, 0, 0], [200, 190, 180], [190, 200, 180], [180, 190, 200], [210, 180, 190],
     [180, 210, 190], [190, 180, 210], [230, 120, 100], [100, 140, 230]], dtype=np.float32)
_MEAN = np.array([179.66761167, 179.55742948, 188.2114891], dtype=np.float32)
_STD = np.array([12.46442902, 12.55030275, 13.12021586], dtype=np.float32)


def scene_sdf(p):
    """Signed distance (positive = free space) and label of the nearest surface.

    p: (...,3) float tensor, world metres.  Returns (sdf, label u8)."""
    ax = p.abs()
    d_walls = ROOM_HALF - ax                                   # (...,3), positive inside
    d_room, wall_axis = d_walls.min(dim=-1)
    wall_side = (torch.gather(p, -1, wall_axis.unsqueeze(-1)).squeeze(-1) > 0).long()
    lab_room = 1 + 2 * wall_axis + wall_side                   # 1..6
    c = torch.tensor(SPHERE_C, dtype=p.dtype, device=p.device)
    d_sph = (p - c).norm(dim=-1) - SPHERE_R
    bc = torch.tensor(BOX_C, dtype=p.dtype, device=p.device)
    bh = torch.tensor(BOX_H, dtype=p.dtype, device=p.device)
    q = (p - bc).abs() - bh
    d_box = q.clamp(min=0).norm(dim=-1) + q.max(dim=-1).values.clamp(max=0)
    sdf = torch.minimum(torch.minimum(d_room, d_sph), d_box)
    lab = lab_room
    lab = torch.where(d_sph <= torch.minimum(d_room, d_box), torch.full_like(lab, LABEL_SPHERE), lab)
    lab = torch.where(d_box <= torch.minimum(d_room, d_sph), torch.full_like(lab, LABEL_BOX), lab)
    return sdf, lab.to(torch.uint8)


def look_at(eye, target, up=(0.0, 0.0, 1.0)):
    """cam->world 4x4 (f32) with camera z forward, y down, x right."""
    eye = np.asarray(eye, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    upv = np.asarray(up, dtype=np.float64)
    right = np.cross(fwd, upv)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    E = np.eye(4)
    E[:3, 0], E[:3, 1], E[:3, 2], E[:3, 3] = right, down, fwd, eye
    return E.astype(np.float32)


def orbit_pose(i, n, radius=0.75, height=0.45, seed=0):
    """Pose i of an n-pose orbit around the room centre, looking across and down into the room.
    The eye stays above the sphere and the cuboid (>= 0.3 m from every surface), like a hand-held
    scan; near-surface frames are exercised by the parity tests, not by the benchmark stream."""
    rng = np.random.RandomState(1911 + 7919 * seed + i)
    a = 2.0 * math.pi * i / max(n, 1) + 0.3 * seed
    eye = (radius * math.cos(a), radius * math.sin(a), height + 0.2 * math.sin(3 * a) + 0.03 * rng.randn())
    tgt = (-0.9 * math.cos(a + 0.4), -0.9 * math.sin(a + 0.4), -0.5 + 0.3 * math.cos(2 * a))
    return look_at(eye, tgt)


def intrinsics_replica(h, w):
    """dataset/replica.py:281-288: f = cx = cy = resolution[0]/2 (90 deg hfov, square-image quirk)."""
    f = h / 2.0
    return np.array([[f, 0.0, f], [0.0, f, f], [0.0, 0.0, 1.0]], dtype=np.float64)


def intrinsics_pinhole(h, w):
    """Centred pinhole, f = w/2."""
    f = w / 2.0
    return np.array([[f, 0.0, w / 2.0], [0.0, f, h / 2.0], [0.0, 0.0, 1.0]], dtype=np.float64)


def render_frame(E, K, h, w, device='cpu', noise=0.005, seed=0, n_steps=96, holes=0.02):
    """Sphere-trace the analytic scene.  Returns depth_gt, noisy depth, mask, labels, image."""
    dev = torch.device(device)
    g = torch.Generator(device='cpu').manual_seed(1911 + seed)
    Ed = torch.from_numpy(np.asarray(E, dtype=np.float64)).to(dev)
    Kinv = torch.from_numpy(np.linalg.inv(np.asarray(K, dtype=np.float64))).to(dev)
    rows, cols = torch.meshgrid(torch.arange(h, dtype=torch.float64, device=dev),
                                torch.arange(w, dtype=torch.float64, device=dev), indexing='ij')
    pix = torch.stack([cols, rows, torch.ones_like(rows)], dim=-1)          # (h,w,3)
    ray_c = pix @ Kinv.T                                                     # z == 1
    ray_w = ray_c @ Ed[:3, :3].T
    scale = ray_w.norm(dim=-1, keepdim=True)
    dir_w = ray_w / scale
    eye = Ed[:3, 3]
    t = torch.zeros(h, w, dtype=torch.float64, device=dev)
    for _ in range(n_steps):
        d, _ = scene_sdf(eye + t.unsqueeze(-1) * dir_w)
        t = t + d.clamp(min=0.0)
    hit = eye + t.unsqueeze(-1) * dir_w
    _, lab = scene_sdf(hit)
    z = (t / scale.squeeze(-1)).float()                                     # z-depth
    nz = torch.randn(h, w, generator=g).to(dev)
    depth = z * (1.0 + noise * nz * z)
    if holes > 0:
        drop = (torch.rand(h, w, generator=g) < holes).to(dev)
        depth = torch.where(drop, torch.zeros_like(depth), depth)
    mask = (depth > 0.05) & (depth < 5.0)
    pal = torch.from_numpy(_PALETTE).to(dev)
    shade = (0.6 + 0.4 * (1.0 / (1.0 + z))).unsqueeze(-1)
    rgb = pal[lab.long()] * shade
    rgb = (rgb - torch.from_numpy(_MEAN).to(dev)) / torch.from_numpy(_STD).to(dev)
    image = rgb.permute(2, 0, 1).contiguous().float()
    return z, depth.float(), mask, lab, image


class SyntheticScene:
    """One analytic scene: a voxel grid description plus a stream of RGB-D frames."""

    def __init__(self, name='synth0', grid=64, h=120, w=160, n_frames=100, seed=0,
                 intrinsics='pinhole', extent=3.2, input_key='tof_depth', pose_rows=4):
        self.name, self.grid, self.h, self.w = name, int(grid), int(h), int(w)
        self.n_frames, self.seed, self.input_key = int(n_frames), int(seed), input_key
        self.resolution = float(extent) / self.grid
        self.origin = np.full(3, -extent / 2.0, dtype=np.float64)
        self.bbox = np.stack([self.origin, self.origin + extent], axis=1)
        self.K = intrinsics_replica(h, w) if intrinsics == 'replica' else intrinsics_pinhole(h, w)
        self.pose_rows = pose_rows

    def gt_volumes(self, device='cpu', truncation=0.1):
        """GT TSDF clamped to +-truncation (dataset/replica.py:305-310) and GT label grid."""
        G = self.grid
        ax = (torch.arange(G, dtype=torch.float64, device=device) + 0.5) * self.resolution
        o = torch.from_numpy(self.origin).to(device)
        sdf = torch.empty(G, G, G, dtype=torch.float16, device=device)
        lab = torch.empty(G, G, G, dtype=torch.uint8, device=device)
        for x0 in range(0, G, 32):
            xs = ax[x0:x0 + 32] + o[0]
            X, Y, Z = torch.meshgrid(xs, ax + o[1], ax + o[2], indexing='ij')
            s, l = scene_sdf(torch.stack([X, Y, Z], dim=-1))
            sc = s.clamp(-truncation, truncation)
            l = torch.where(s.abs() > truncation, torch.zeros_like(l), l)
            sdf[x0:x0 + 32] = sc.to(torch.float16)
            lab[x0:x0 + 32] = l
        return sdf, lab

    def frame(self, i, device='cpu'):
        """Sample dict for frame i, batched with a leading 1 like a DataLoader(batch_size=1)."""
        E = orbit_pose(i, self.n_frames, seed=self.seed)
        z, depth, mask, lab, image = render_frame(E, self.K, self.h, self.w, device=device,
                                                  seed=self.seed * 100003 + i)
        ext = torch.from_numpy(E[:self.pose_rows]).to(device)
        return {
            'image': image.unsqueeze(0),
            self.input_key: depth.unsqueeze(0),
            'depth_gt': z.unsqueeze(0),
            'mask': mask.unsqueeze(0),
            'extrinsics': ext.unsqueeze(0),
            'intrinsics': torch.from_numpy(self.K).to(device).unsqueeze(0),
            'semantic_gt': lab.unsqueeze(0),
            'frame_id': ['%s/0/%d' % (self.name, i)],
        }


def seeded_parameters(module, seed, scale=1.0):
    """Fill every parameter and buffer of `module` from a seeded generator, in sorted state_dict key order, so that two
    modules with the same key set (the reference's network and this package's mirror) get bit-identical values no
    matter in which order their constructors created the tensors.  Weights ~ N(0, fan-in scaled), BatchNorm statistics
    non-trivial (running_var in [0.5, 1.5]), like a trained checkpoint has.  No checkpoint exists offline."""
    gen = torch.Generator().manual_seed(int(seed))
    sd = module.state_dict()
    with torch.no_grad():
        for k in sorted(sd):
            t = sd[k]
            if not t.is_floating_point():
                continue                                   # num_batches_tracked
            if k.endswith('running_var'):
                v = 0.5 + torch.rand(t.shape, generator=gen)
            elif k.endswith('running_mean'):
                v = 0.1 * torch.randn(t.shape, generator=gen)
            elif t.dim() >= 2:
                fan_in = t[0].numel()
                v = torch.randn(t.shape, generator=gen) * (scale * (2.0 / fan_in) ** 0.5)
            elif k.endswith('weight'):
                v = 0.8 + 0.4 * torch.rand(t.shape, generator=gen)     # BatchNorm gamma
            else:
                v = 0.05 * torch.randn(t.shape, generator=gen)         # biases / BatchNorm beta
            t.copy_(v.to(t.dtype))
    return module
