"""Pipeline.fuse / fuse_training on the GPU: the plumbing around the two kernels.

The network output of each frame is captured and fed to the CPU oracle, so the volumes must be
bit-exact regardless of how the (floating-point) convolutions were computed."""
import numpy as np
import pytest
import torch

from oracle import oracle
from online_joint_depthfusion_and_semantic_b200.config import Config, fusion_config
from online_joint_depthfusion_and_semantic_b200.modules.database import Database, Voxelgrid
from online_joint_depthfusion_and_semantic_b200.modules.pipeline import Pipeline
from online_joint_depthfusion_and_semantic_b200.synthetic import SyntheticScene

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


class _Set:
    def __init__(self, scene):
        self.s, self.scenes = scene, [scene.name]

    def get_grid(self, name, truncation, semantic_grid):
        sdf, lab = self.s.gt_volumes(device=DEV, truncation=truncation)
        g = Voxelgrid(self.s.resolution); g.from_array(sdf, self.s.bbox)
        l = Voxelgrid(self.s.resolution); l.from_array(lab, self.s.bbox)
        return (g, l)


def _world(h, w, G, strategy, use_sem, semantics='class30'):
    torch.manual_seed(1911)
    scene = SyntheticScene(name='s0', grid=G, h=h, w=w, n_frames=6, seed=5)
    cfg = fusion_config(h, w, semantics=semantics, semantic_strategy=strategy, use_semantics=use_sem, device='cuda:0')
    pipe = Pipeline(cfg).to(DEV).eval()
    if pipe._semantic_2d_network is not None:
        pipe._semantic_2d_network.set_bottleneck_dropout(False)
    db = Database(_Set(scene), Config(device=DEV, implementation='efficient', init_value=0.1, semantics=semantics,
                                      semantic_grid=bool(semantics)))
    return scene, cfg, pipe, db


@pytest.mark.parametrize('strategy,use_sem', [('gt', False), ('predict', True)])
def test_fuse_volumes_bit_exact_given_network_output(strategy, use_sem):
    h, w, G = 48, 64, 48
    scene, cfg, pipe, db = _world(h, w, G, strategy, use_sem)
    captured = {}
    inner = pipe._fusion

    def tap(inputs, values, **kw):
        est = inner(inputs, values, **kw)
        captured['est'] = est.detach()[0].cpu().numpy().copy()
        captured['world'] = values['pcl'][0].cpu().numpy().copy()
        captured['vals'] = values['fusion_values'][0].cpu().numpy().copy()
        return est
    pipe._fusion = tap
    seg = {}
    if strategy == 'predict':
        inner_seg = pipe._semantic_frame

        def seg_tap(batch, as_uint8):
            sc, ids = inner_seg(batch, as_uint8)
            seg['scores'], seg['ids'] = sc.reshape(-1).cpu().numpy().copy(), ids.reshape(-1).to(torch.uint8).cpu().numpy().copy()
            return sc, ids
        pipe._semantic_frame = seg_tap
    t_o = np.full((G, G, G), 0.1, np.float16).view(np.uint16)
    w_o = np.zeros((G, G, G), np.uint16); i_o = np.zeros((G, G, G), np.uint8); s_o = np.zeros((G, G, G), np.uint16)
    with torch.no_grad():
        for i in range(3):
            b = scene.frame(i, device=DEV)
            depth, mask = b['tof_depth'][0].cpu().numpy(), b['mask'][0].cpu().numpy()
            E = b['extrinsics'][0].cpu().numpy()
            pipe.fuse(b, db, DEV)
            torch.cuda.synchronize()
            o = oracle.extract(captured['world'], E[:3, 3], scene.origin, scene.resolution, t_o, w_o)
            assert np.array_equal(o['fusion_values'].view(np.uint32), captured['vals'].view(np.uint32))
            filt = np.where(mask, depth, np.float32(0)).reshape(-1)
            ids = seg['ids'] if strategy == 'predict' else b['semantic_gt'].reshape(-1).cpu().numpy()
            sc = seg['scores'] if strategy == 'predict' else np.ones(h * w, np.float32)
            oracle.integrate_frame(captured['world'], filt, captured['est'], E[:3, 3], scene.origin, scene.resolution,
                                   t_o, w_o, pix_ids=ids, pix_scores=sc, ids_vol=i_o, scores_vol=s_o, do_sem=True)
    assert db.state['s0'] is True
    assert np.array_equal(db.scenes_est['s0'].volume.cpu().numpy().view(np.uint16), t_o)
    assert np.array_equal(db.fusion_weights['s0'].cpu().numpy().view(np.uint16), w_o)
    assert np.array_equal(db.ids_est['s0'].volume.cpu().numpy(), i_o)
    assert np.array_equal(db.scores['s0'].volume.cpu().numpy().view(np.uint16), s_o)
    assert int((w_o != 0).sum()) > 1000


def test_fuse_training_outputs_and_no_semantic_update():
    h, w, G = 48, 64, 48
    scene, cfg, pipe, db = _world(h, w, G, 'gt', True)
    pipe._fusion_network.train()
    b = scene.frame(0, device=DEV)
    out = pipe.fuse_training(b, db, DEV)
    nv = int((b['mask'] & (b['tof_depth'] != 0)).sum())
    assert out['tsdf_est'].shape == (1, h * w, 9)
    assert out['tsdf_fused'].shape == (1, nv, 9) and out['tsdf_target'].shape == (1, nv, 9)
    assert out['tsdf_fused'].requires_grad and not out['tsdf_target'].requires_grad
    loss = (out['tsdf_fused'] - out['tsdf_target']).abs().mean()
    loss.backward()
    g = [p.grad for p in pipe._fusion_network.parameters() if p.grad is not None]
    assert len(g) > 100 and all(torch.isfinite(x).all() for x in g)
    assert bool((db.ids_est['s0'].volume == 0).all()) and bool((db.scores['s0'].volume == 0).all())   # test=False
    assert int((db.fusion_weights['s0'] > 0).sum()) > 1000
    assert not db.scenes_est['s0'].volume.requires_grad
    # the target is the trilinear read of the GT volume at the same samples
    tgt = out['tsdf_target']
    assert float(tgt.abs().max()) <= 0.1 + 1e-6


def test_no_semantics_config():
    h, w, G = 48, 64, 32
    scene, cfg, pipe, db = _world(h, w, G, 'gt', False, semantics='')
    assert pipe._semantic_2d_network is None
    with torch.no_grad():
        pipe.fuse(scene.frame(0, device=DEV), db, DEV)
    torch.cuda.synchronize()
    assert int((db.fusion_weights['s0'] > 0).sum()) > 500


class _ClassicUpdate(torch.nn.Module):
    """The stand-in tests/golden/make_golden.py puts in place of FusionNet: est[n,k] = (4-k)*resolution."""

    def __init__(self, res, n_points=9):
        super().__init__()
        self.res, self.n_points = res, n_points

    def forward(self, x):
        b, _, h, w = x['tsdf_values'].shape
        k = torch.arange(self.n_points, dtype=torch.float32, device=x['tsdf_values'].device)
        return (((self.n_points // 2) - k) * self.res).view(1, -1, 1, 1).expand(b, -1, h, w).contiguous()


def test_fuse_training_values_vs_reference_golden(golden):
    """Row a2 / a12: `Pipeline.fuse_training` against the REFERENCE's own fuse_training (fixture training_24x32_g32):
    tsdf_est, tsdf_fused (running-mean formula + masking, modules/pipeline.py:104-135,365-405), tsdf_target (second
    extraction from the GT volume, :309-315) of every frame, and the volumes after the test=False integration --
    bit for bit, given the reference's own world points."""
    import numpy as np
    from online_joint_depthfusion_and_semantic_b200.modules.database import Database, Voxelgrid
    from online_joint_depthfusion_and_semantic_b200.config import Config
    g = golden('training_24x32_g32')
    G, res, h, w = int(g['G']), float(g['res']), 24, 32
    cfg = fusion_config(h, w, semantic_strategy='gt', use_semantics=False)
    pipe = Pipeline(cfg).to(DEV)
    pipe._fusion_network = _ClassicUpdate(res)
    pipe.train()

    class DS:
        scenes = ['synth0']

        def get_grid(self, s, init, sem):
            vg = Voxelgrid(res)
            vg.from_array(torch.from_numpy(g['gt'].view(np.float16).copy()), np.stack([g['origin'], g['origin'] + G * res], 1))
            lab = Voxelgrid(res)
            lab.from_array(torch.zeros(G, G, G, dtype=torch.uint8), vg.bbox)
            return vg, lab
    db = Database(DS(), Config(device=DEV, implementation='efficient', init_value=0.1, semantics='class30',
                               semantic_grid=True, n_classes=30))
    inner = pipe._extractor.forward
    for j in range(int(g['n_frames'])):
        pcl = torch.from_numpy(g['f%d_pcl' % j]).to(DEV)
        pipe._extractor.forward = lambda *a, _p=pcl, **k: inner(*a, world=_p, **k)       # the reference's BLAS-ordered points
        batch = {'image': torch.zeros(1, 3, h, w, device=DEV), 'tof_depth': torch.from_numpy(g['f%d_tof_depth' % j]).to(DEV),
                 'mask': torch.from_numpy(g['f%d_mask' % j]).to(DEV), 'extrinsics': torch.from_numpy(g['f%d_extrinsics' % j]),
                 'intrinsics': torch.from_numpy(g['f%d_intrinsics' % j]),
                 'semantic_gt': torch.from_numpy(g['f%d_semantic_gt' % j]).to(DEV), 'frame_id': ['synth0/0/%d' % j]}
        out = pipe.fuse_training(batch, db, DEV)
        for k in ('tsdf_est', 'tsdf_fused', 'tsdf_target'):
            got, want = out[k].detach().cpu().numpy(), g['f%d_%s' % (j, k)]
            assert got.shape == want.shape, (j, k, got.shape, want.shape)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (j, k, float(np.abs(got - want).max()))
    torch.cuda.synchronize()
    assert np.array_equal(db.scenes_est['synth0'].volume.cpu().numpy().view(np.uint16), g['tsdf'])
    assert np.array_equal(db.fusion_weights['synth0'].cpu().numpy().view(np.uint16), g['wvol'])
    assert int(db.ids_est['synth0'].volume.count_nonzero()) == 0             # test=False: no semantic update


@pytest.mark.parametrize('strategy,use_sem', [('gt', False), ('predict', True)])
def test_frame_stream_equals_synchronous_fuse(strategy, use_sem):
    """stream.FrameStream (pinned host frames, copies / read-backs overlapped with the kernels of the neighbouring frames;
    with predicted semantics also the AdapNet++ pass of frame i+1 on its own stream next to the fusion of frame i)
    leaves bit-identical volumes to calling Pipeline.fuse frame by frame, and returns every frame's result in order."""
    from online_joint_depthfusion_and_semantic_b200.stream import FrameStream
    h, w, G = 48, 64, 48
    vols, results = [], []
    for streamed in (False, True):
        scene, cfg, pipe, db = _world(h, w, G, strategy, use_sem)
        last = {}
        inner = pipe._fusion

        def tap(inputs, values, _inner=inner, _last=last, **kw):
            est = _inner(inputs, values, **kw)
            _last['v'] = est.abs().mean()
            return est
        pipe._fusion = tap
        frames = [{k: (v.cpu().pin_memory() if torch.is_tensor(v) else v) for k, v in scene.frame(i, device='cpu').items()} for i in range(5)]
        got = []
        with torch.no_grad():
            if streamed:
                fs = FrameStream(pipe, db, DEV, result_fn=lambda: last['v'], depth=2)
                for b in frames:
                    r = fs.submit(b)
                    if r is not None:
                        got.append(r)
                got += fs.flush()
                assert fs.h2d_bytes > 0
            else:
                for b in frames:
                    pipe.fuse({k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in b.items()}, db, DEV)
                    got.append(float(last['v'].item()))
        torch.cuda.synchronize()
        v = db['s0']
        vols.append([v['current'].cpu().numpy().view(np.uint16).copy(), v['weights'].cpu().numpy().view(np.uint16).copy()] +
                    ([v['ids_est'].cpu().numpy().copy(), v['scores'].cpu().numpy().view(np.uint16).copy()] if use_sem else []))
        results.append(got)
    assert len(results[1]) == 5 and results[0] == results[1]
    for a, b in zip(vols[0], vols[1]):
        assert np.array_equal(a, b)
