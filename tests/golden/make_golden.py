"""Generate the golden fixtures in this directory from the REFERENCE ITSELF.

Run only in the build container (needs /root/reference; the GPU box does not have it):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports the reference's own modules/extractor.py, modules/integrator.py and
modules/pipeline.py (suryanshkumar/online-joint-depthfusion-and-semantic @ a4f9e19)
on CPU with ONE torch thread (the semantic scatter of modules/integrator.py:123-124 is
only deterministic single-threaded), feeds them seeded inputs and
stores inputs + outputs as small npz files.  tests/test_oracle_golden.py pins
oracle/ojdf_oracle.c to these files; the GPU parity tests then compare the CUDA path
with that oracle and with these files directly.

fp16 volumes are stored as their uint16 bit patterns; voxel indices as int16.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get('OJDF_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def ref_env():
    """Make the reference importable here (SURVEY.md App. D recipe)."""
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ('matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(name, types.ModuleType(name))
    import torchvision
    import modules.adapnet as ref_adapnet
    ref_adapnet.resnet50 = lambda pretrained=True: torchvision.models.resnet50(weights=None)
    import modules.extractor as ref_extractor
    import modules.integrator as ref_integrator
    import modules.pipeline as ref_pipeline
    import modules.model as ref_model
    return ref_extractor, ref_integrator, ref_pipeline, ref_model, ref_adapnet


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def make_config(h, w, semantics=True, strategy='gt', use_semantics=False, n_classes=30):
    return AttrDict(
        SETTINGS=AttrDict(gpu=False, device=torch.device('cpu'), implementation='efficient'),
        FUSION_MODEL=AttrDict(name='v3', output_scale=1.0, n_points=9, n_tail_points=7, growth_factor=6,
                              use_semantics=use_semantics),
        SEMANTIC_2D_MODEL=AttrDict(stage=2, n_classes=n_classes),
        DATA=AttrDict(semantics='class30' if semantics else '', semantic_strategy=strategy, input='tof_depth',
                      resx=w, resy=h, init_value=0.1),
    )


def h16(a):
    return np.ascontiguousarray(a).view(np.uint16)


def rand_pose(rs, rows=4, eye=None):
    """Random rotation (QR) + eye; cam->world."""
    q, _ = np.linalg.qr(rs.randn(3, 3))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    E = np.eye(4)
    E[:3, :3] = q
    E[:3, 3] = eye if eye is not None else rs.uniform(-0.3, 0.3, 3)
    return E[:rows].astype(np.float32)


def rand_volumes(rs, G, zero_weight_frac=0.5):
    tsdf = rs.uniform(-0.1, 0.1, (G, G, G)).astype(np.float16)
    wv = rs.uniform(0.0, 20.0, (G, G, G)).astype(np.float16)
    wv[rs.rand(G, G, G) < zero_weight_frac] = 0
    return tsdf, wv


def case_extract(name, h, w, G, seed, rows, ext, depth_lo, depth_hi, store_full=True):
    ref_extractor, *_ = ref_env()
    rs = np.random.RandomState(seed)
    depth = rs.uniform(depth_lo, depth_hi, (1, h, w)).astype(np.float32)
    depth[0, rs.rand(h, w) < 0.05] = 0.0                       # holes: world == eye, direction 0
    f = w / 2.0
    K = np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1]], dtype=np.float64)
    E = rand_pose(rs, rows)
    res = ext / G
    origin = np.full(3, -ext / 2.0) + rs.uniform(-0.01, 0.01, 3)
    tsdf, wv = rand_volumes(rs, G)
    cfg = make_config(h, w)
    ex = ref_extractor.Extractor(cfg)
    out = ex.forward(torch.from_numpy(depth), torch.from_numpy(E[None]), torch.from_numpy(K[None]),
                     torch.from_numpy(tsdf), torch.from_numpy(wv), torch.from_numpy(origin), res)
    assert out['points'].dtype == torch.float64 and out['weights'].dtype == torch.float64
    assert out['indices'].dtype == torch.int64 and out['fusion_values'].dtype == torch.float32
    Kinv = torch.from_numpy(K).float().inverse().numpy()
    idx = out['indices'].numpy()
    assert idx.min() > -30000 and idx.max() < 30000
    d = dict(depth=depth, E=E, K=K, Kinv=Kinv, origin=origin, res=np.float64(res),
             tsdf=h16(tsdf), wvol=h16(wv), G=np.int32(G),
             pcl=out['pcl'].numpy(), fusion_values=out['fusion_values'].numpy(),
             fusion_weights=out['fusion_weights'].numpy())
    full = dict(points=out['points'].numpy(), indices=idx.astype(np.int16), weights=out['weights'].numpy())
    if store_full:
        d.update(full)
    else:
        for k, v in full.items():
            d['sha256_' + k] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(v).tobytes()).digest(), dtype=np.uint8)
    oob = float(((idx < 0) | (idx >= G)).any(-1).mean())
    print('%-22s N=%d G=%d oob corners %.1f%%' % (name, h * w, G, 100 * oob))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **d)
    return d, out


def case_integrate(name, h, w, G, seed, ext, do_sem=True, test=True):
    """Reference Integrator.forward on the reference Extractor's own indices/weights."""
    ref_extractor, ref_integrator, *_ = ref_env()
    rs = np.random.RandomState(seed)
    depth = rs.uniform(0.4, 1.6, (1, h, w)).astype(np.float32)
    f = w / 2.0
    K = np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1]], dtype=np.float64)
    E = rand_pose(rs, 4)
    res = ext / G
    origin = np.full(3, -ext / 2.0)
    tsdf, wv = rand_volumes(rs, G)
    ids_vol = rs.randint(0, 6, (G, G, G)).astype(np.uint8)
    sc_vol = (rs.randint(0, 5, (G, G, G)) / 4.0).astype(np.float16)          # many exact ties
    cfg = make_config(h, w, semantics=do_sem)
    ex = ref_extractor.Extractor(cfg)
    out = ex.forward(torch.from_numpy(depth), torch.from_numpy(E[None]), torch.from_numpy(K[None]),
                     torch.from_numpy(tsdf), torch.from_numpy(wv), torch.from_numpy(origin), res)
    N = h * w
    filt = depth.reshape(N).copy()
    filt[rs.rand(N) < 0.2] = 0.0                                                # masked rays
    valid = np.nonzero(filt != 0)[0]
    est = rs.uniform(-0.15, 0.15, (N, 9)).astype(np.float32)
    pix_ids = rs.randint(0, 6, N).astype(np.uint8)
    pix_sc = (rs.randint(0, 5, N) / 4.0).astype(np.float32)
    T = 7
    vt = torch.from_numpy(valid)
    upd = dict(values=torch.clamp(torch.from_numpy(est)[None][:, vt, :T], -0.1, 0.1),
               indices=out['indices'][:, vt, :T], weights=out['weights'][:, vt, :T],
               points=out['points'][:, vt, :T],
               semantics=torch.from_numpy(pix_ids)[None, :, None, None].repeat(1, 1, 9, 1)[:, vt, :T],
               scores=torch.from_numpy(pix_sc)[None, :, None, None].repeat(1, 1, 9, 1)[:, vt, :T])
    integ = ref_integrator.Integrator(cfg)
    t0, w0, i0, s0 = (torch.from_numpy(a.copy()) for a in (tsdf, wv, ids_vol, sc_vol))
    v1, w1, i1, s1 = integ.forward(upd, t0, w0, s0, i0, test=test)
    d = dict(depth=depth, filt=filt, E=E, K=K, Kinv=torch.from_numpy(K).float().inverse().numpy(),
             origin=origin, res=np.float64(res), G=np.int32(G), est=est, pix_ids=pix_ids, pix_scores=pix_sc,
             pcl=out['pcl'].numpy(), do_sem=np.int32(do_sem and test),
             tsdf0=h16(tsdf), wvol0=h16(wv), ids0=ids_vol, scores0=h16(sc_vol),
             tsdf1=h16(v1.numpy()), wvol1=h16(w1.numpy()), ids1=i1.numpy(), scores1=h16(s1.numpy()))
    touched = int((h16(w1.numpy()) != h16(wv)).sum())
    nan = int(np.isnan(v1.numpy().astype(np.float32)).sum())
    print('%-22s Nv=%d touched(weight changed)=%d NaN voxels=%d labels changed=%d' %
          (name, len(valid), touched, nan, int((i1.numpy() != ids_vol).sum())))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **d)


def case_integrate_nan(name, G=16, seed=31):
    """0/0 -> NaN (modules/integrator.py:82): the reference Integrator.forward on a hand-made `updates` dict in which
    some voxels with a zero prior weight only ever receive zero-weight entries (in real frames: corner weights that
    underflow in the f64 -> f32 cast, SURVEY.md section 7 "0/0 NaN").  Also has voxels hit by several entries of which
    only some are zero, entries outside the grid, and a voxel whose prior weight is non-zero but whose entries are all
    zero (stays finite)."""
    _, ref_integrator, *_ = ref_env()
    rs = np.random.RandomState(seed)
    Nv, T = 40, 7
    idx = rs.randint(-2, G + 2, (1, Nv, T, 8, 3)).astype(np.int64)
    wts = rs.uniform(0.0, 1.0, (1, Nv, T, 8))
    wts[rs.rand(1, Nv, T, 8) < 0.35] = 0.0
    vals = rs.uniform(-0.1, 0.1, (1, Nv, T)).astype(np.float32)
    tsdf, wv = rand_volumes(rs, G, zero_weight_frac=0.6)
    # make sure the interesting cases exist: a fresh voxel with only zero-weight entries, and one with prior weight
    wv[3, 4, 5], wv[6, 7, 8] = 0.0, 2.5
    for vx in ((3, 4, 5), (6, 7, 8)):                          # nobody else touches the two special voxels
        hit = (idx == np.array(vx)).all(-1)
        idx[hit] = -1
    for n, vx in enumerate(((3, 4, 5), (6, 7, 8))):
        idx[0, n, :3, :2] = vx
        wts[0, n, :3, :2] = 0.0
    ids_vol = rs.randint(0, 6, (G, G, G)).astype(np.uint8)
    sc_vol = (rs.randint(0, 5, (G, G, G)) / 4.0).astype(np.float16)
    pix_ids = rs.randint(0, 6, (1, Nv, T, 1)).astype(np.uint8)
    pix_sc = (rs.randint(0, 5, (1, Nv, T, 1)) / 4.0).astype(np.float32)
    cfg = make_config(8, 8, semantics=True)
    upd = dict(values=torch.from_numpy(vals), indices=torch.from_numpy(idx), weights=torch.from_numpy(wts),
               points=torch.zeros(1, Nv, T, 3, dtype=torch.float64), semantics=torch.from_numpy(pix_ids),
               scores=torch.from_numpy(pix_sc))
    t0, w0, i0, s0 = (torch.from_numpy(a.copy()) for a in (tsdf, wv, ids_vol, sc_vol))
    v1, w1, i1, s1 = ref_integrator.Integrator(cfg).forward(upd, t0, w0, s0, i0, test=True)
    nan = int(np.isnan(v1.numpy().astype(np.float32)).sum())
    assert nan > 0 and np.isnan(np.float32(v1[3, 4, 5])) and np.isfinite(np.float32(v1[6, 7, 8]))
    print('%-22s entries=%d NaN voxels=%d' % (name, Nv * T * 8, nan))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), values=vals, indices=idx.astype(np.int16), weights=wts,
                        semantics=pix_ids, scores=pix_sc, G=np.int32(G),
                        tsdf0=h16(tsdf), wvol0=h16(wv), ids0=ids_vol, scores0=h16(sc_vol),
                        tsdf1=h16(v1.numpy()), wvol1=h16(w1.numpy()), ids1=i1.numpy(), scores1=h16(s1.numpy()))


class ClassicUpdate(torch.nn.Module):
    """Stand-in for FusionNet inside the REFERENCE Pipeline: the classical TSDF update
    est[n,k] = (4-k)*resolution, so that the multi-frame
    fixture pins Extractor + _prepare_volume_update + Integrator without any conv net."""

    def __init__(self, res, n_points=9):
        super().__init__()
        self.res, self.n_points = res, n_points

    def forward(self, x):
        b, _, h, w = x['tsdf_values'].shape
        k = torch.arange(self.n_points, dtype=torch.float32)
        prof = ((self.n_points // 2) - k) * self.res
        return prof.view(1, -1, 1, 1).expand(b, -1, h, w).contiguous()


class FakeGrid:
    def __init__(self, volume):
        self.volume = volume


class FakeDatabase:
    """What modules/pipeline.py:199-244 touches of modules/database.py."""

    def __init__(self, name, G, origin, res, init_value=0.1):
        self.state = {name: False}
        self.origin = {name: torch.from_numpy(np.asarray(origin, dtype=np.float64))}
        self.resolution = {name: float(res)}
        self.scenes_est = {name: FakeGrid(torch.full((G, G, G), init_value, dtype=torch.float16))}
        self.scenes_gt = {name: FakeGrid(torch.zeros((G, G, G), dtype=torch.float16))}
        self.fusion_weights = {name: torch.zeros((G, G, G), dtype=torch.float16)}
        self.ids_est = {name: FakeGrid(torch.zeros((G, G, G), dtype=torch.uint8))}
        self.scores = {name: FakeGrid(torch.zeros((G, G, G), dtype=torch.float16))}

    def __getitem__(self, s):
        return dict(origin=self.origin[s], resolution=self.resolution[s], gt=self.scenes_gt[s].volume,
                    current=self.scenes_est[s].volume, weights=self.fusion_weights[s],
                    ids_est=self.ids_est[s].volume, scores=self.scores[s].volume)


def case_pipeline(name, h, w, G, n_frames, frame_ids):
    """Reference Pipeline.fuse over a short synthetic orbit with gt semantics."""
    *_, ref_pipeline, _, _ = ref_env()
    from online_joint_depthfusion_and_semantic_b200.synthetic import SyntheticScene
    scene = SyntheticScene(name='synth0', grid=G, h=h, w=w, n_frames=n_frames, seed=3)
    cfg = make_config(h, w, semantics=True, strategy='gt')
    pipe = ref_pipeline.Pipeline(cfg)
    pipe._fusion_network = ClassicUpdate(scene.resolution)
    pipe.eval()
    db = FakeDatabase('synth0', G, scene.origin, scene.resolution)
    frames = {}
    pcls = []
    ref_forward = pipe._extractor.forward

    def recording_forward(*a, **k):          # keep the reference's own BLAS-ordered world points
        out = ref_forward(*a, **k)
        pcls.append(out['pcl'].numpy().copy())
        return out
    pipe._extractor.forward = recording_forward
    with torch.no_grad():
        for j, i in enumerate(frame_ids):
            b = scene.frame(i)
            for k in ('tof_depth', 'mask', 'extrinsics', 'intrinsics', 'semantic_gt'):
                frames['f%d_%s' % (j, k)] = b[k].numpy().copy()
            b['image'] = torch.zeros(1, 3, h, w)
            pipe.fuse(b, db, torch.device('cpu'))
            frames['f%d_pcl' % j] = pcls[-1]
    d = dict(frames, n_frames=np.int32(len(frame_ids)), G=np.int32(G), origin=scene.origin,
             res=np.float64(scene.resolution),
             tsdf=h16(db.scenes_est['synth0'].volume.numpy()), wvol=h16(db.fusion_weights['synth0'].numpy()),
             ids=db.ids_est['synth0'].volume.numpy(), scores=h16(db.scores['synth0'].volume.numpy()))
    wv = db.fusion_weights['synth0'].numpy().astype(np.float32)
    print('%-22s frames=%d voxels with weight>0: %d  labelled: %d' %
          (name, len(frame_ids), int((wv > 0).sum()), int((d['ids'] > 0).sum())))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **d)


def case_training(name, h, w, G, n_frames, frame_ids):
    """Reference Pipeline.fuse_training (modules/pipeline.py:251-363) over a few frames: the loss tensors
    tsdf_est / tsdf_fused / tsdf_target of every frame (incl. the reference's _prepare_fusion_output + masking,
    :104-135,365-405, and the second extraction from the GT volume) and the volumes after the `test=False` integration.
    FusionNet is replaced by the classical update so no convolution is involved; gt semantics."""
    *_, ref_pipeline, _, _ = ref_env()
    from online_joint_depthfusion_and_semantic_b200.synthetic import SyntheticScene
    scene = SyntheticScene(name='synth0', grid=G, h=h, w=w, n_frames=n_frames, seed=5)
    cfg = make_config(h, w, semantics=True, strategy='gt')
    pipe = ref_pipeline.Pipeline(cfg)
    pipe._fusion_network = ClassicUpdate(scene.resolution)
    pipe.train()
    db = FakeDatabase('synth0', G, scene.origin, scene.resolution)
    gt, _ = scene.gt_volumes()
    db.scenes_gt['synth0'].volume = gt
    d = dict(n_frames=np.int32(len(frame_ids)), G=np.int32(G), origin=scene.origin, res=np.float64(scene.resolution),
             gt=h16(gt.numpy()))
    pcls = []
    ref_forward = pipe._extractor.forward

    def recording_forward(*a, **k):
        out = ref_forward(*a, **k)
        pcls.append(out['pcl'].numpy().copy())
        return out
    pipe._extractor.forward = recording_forward
    for j, i in enumerate(frame_ids):
        b = scene.frame(i)
        for k in ('tof_depth', 'mask', 'extrinsics', 'intrinsics', 'semantic_gt'):
            d['f%d_%s' % (j, k)] = b[k].numpy().copy()
        b['image'] = torch.zeros(1, 3, h, w)
        out = pipe.fuse_training(b, db, torch.device('cpu'))
        d['f%d_pcl' % j] = pcls[-1]
        for k in ('tsdf_est', 'tsdf_fused', 'tsdf_target'):
            d['f%d_%s' % (j, k)] = out[k].detach().numpy().copy()
    d.update(tsdf=h16(db.scenes_est['synth0'].volume.numpy()), wvol=h16(db.fusion_weights['synth0'].numpy()),
             ids=db.ids_est['synth0'].volume.numpy(), scores=h16(db.scores['synth0'].volume.numpy()))
    assert int((d['ids'] != 0).sum()) == 0                                   # test=False: the semantic volumes stay untouched
    print('%-22s frames=%d Nv(last)=%d' % (name, len(frame_ids), out['tsdf_fused'].shape[1]))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **d)


def main():
    torch.set_num_threads(1)
    torch.manual_seed(1911)
    # full arrays stored: 24x32 rays, 32^3 grid, ~30 % OOB, 4x4 and 3x4 poses
    case_extract('extract_24x32_g32_a', 24, 32, 32, seed=11, rows=4, ext=2.4, depth_lo=0.3, depth_hi=1.8)
    case_extract('extract_24x32_g32_b', 24, 32, 32, seed=12, rows=3, ext=1.2, depth_lo=0.2, depth_hi=1.5)
    # digests only (indices/weights/points) at the plumbing size of BASELINE.json configs[0]
    case_extract('extract_120x160_g64', 120, 160, 64, seed=13, rows=4, ext=3.2, depth_lo=0.3, depth_hi=2.2,
                 store_full=False)
    case_integrate('integrate_24x32_g32_sem', 24, 32, 32, seed=21, ext=2.4, do_sem=True, test=True)
    case_integrate('integrate_24x32_g32_train', 24, 32, 32, seed=22, ext=2.4, do_sem=True, test=False)
    case_integrate('integrate_48x64_g24_dup', 48, 64, 24, seed=23, ext=3.0, do_sem=True, test=True)
    case_integrate_nan('integrate_updates_nan_g16')
    case_pipeline('pipeline_48x64_g48', 48, 64, 48, n_frames=12, frame_ids=[0, 1, 2, 5])
    case_training('training_24x32_g32', 24, 32, 32, n_frames=12, frame_ids=[0, 1, 4])


if __name__ == '__main__':
    main()
