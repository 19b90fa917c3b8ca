"""Run the UNMODIFIED reference (baseline/_ref/, or /root/reference in the build container) in this image.

The reference imports packages the image does not have (easydict, h5py, trimesh, scikit-image, matplotlib, its own
`graphics` Voxelgrid) and loads a dataset from disk; this module supplies the minimal stand-ins SURVEY.md App. D lists
and a synthetic dataset class that is injected into `utils.setup` under the name the config asks for
(`utils/setup.py:73-77` does `eval(dataset)(config)`).  Nothing here changes a line of the reference: its
`test_fusion.py::test_fusion(config)`, `utils/setup.py`, `modules/database.py`, `utils/metrics.py` run as they are.

Two uses:
  * `drive(impl='reference')`  -- the reference end to end (its own Pipeline / Extractor / Integrator on torch): the
    CPU baseline of bench.py (`kind: "reference"`) and the yardstick of the drive-through parity test;
  * `drive(impl='ours')`       -- INTEGRATION.md option A: `modules.pipeline / extractor / integrator` of the
    reference's import namespace are replaced by this repository's modules before `test_fusion.py` is imported, the
    reference's driver, Database and metrics then run unchanged on top of the CUDA path.

CLI (each run in its own process -- the two variants cannot share sys.modules):
    python baseline/harness.py drive --impl ours|reference --gpu 0|1 [--frames N --h H --w W --grid G --out file.json]
"""
import argparse
import json
import os
import sys
import tempfile
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def ref_root():
    """Where the unmodified reference lives: $OJDF_REFERENCE, /root/reference (build container), baseline/_ref."""
    for cand in (os.environ.get('OJDF_REFERENCE'), '/root/reference', os.path.join(HERE, '_ref')):
        if cand and os.path.isdir(os.path.join(cand, 'modules')):
            return cand
    return None


class EasyDict(dict):
    """easydict.EasyDict as far as the reference uses it: recursive attribute access, AttributeError on a miss."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = __setitem__


class Voxelgrid:
    """deps/graphics Voxelgrid as far as the fusion path touches it (voxelgrid.py:54-70,157-161,196-218)."""

    def __init__(self, resolution):
        self.resolution = resolution
        self._volume = None
        self.bbox = None
        self.origin = None

    def from_array(self, array, bbox):
        self._volume = array
        self.bbox = np.asarray(bbox)
        self.origin = self.bbox[:, 0].copy()

    @property
    def volume(self):
        return self._volume

    @volume.setter
    def volume(self, v):
        self._volume = v

    @property
    def shape(self):
        return tuple(self._volume.shape)


def install_shims():
    """sys.modules stand-ins for what the reference imports and this image lacks (never used on the fusion path)."""
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    def absent(name):
        try:
            __import__(name)
            return False
        except ImportError:
            return True

    if absent('matplotlib'):
        mod('matplotlib').pyplot = mod('matplotlib.pyplot')
    if absent('easydict'):
        mod('easydict', EasyDict=EasyDict)
    for name in ('h5py', 'trimesh', 'plyfile'):
        if absent(name):
            mod(name, PlyElement=None, PlyData=None)
    if absent('skimage'):
        sk = mod('skimage')
        sk.measure = mod('skimage.measure')
        sk.io = mod('skimage.io')
        sk.exposure = mod('skimage.exposure', rescale_intensity=None, is_low_contrast=None)
    if absent('graphics'):
        mod('graphics', Voxelgrid=Voxelgrid)


def enter(root=None):
    """Make the reference importable (its root first on sys.path) and patch the one network download."""
    root = root or ref_root()
    if root is None:
        raise RuntimeError('no reference tree: run `python baseline/install_ref.py` in the build container')
    sys.dont_write_bytecode = True                     # /root/reference is read-only
    if root not in sys.path:
        sys.path.insert(0, root)
    install_shims()
    import torchvision
    import modules.adapnet as ref_adapnet              # noqa: E402  (the reference's own package `modules`)
    ref_adapnet.resnet50 = lambda pretrained=True: torchvision.models.resnet50(weights=None)   # modules/adapnet.py:101
    return root


def swap_in_ours():
    """INTEGRATION.md option A: the reference's import names resolve to this repository's modules."""
    import online_joint_depthfusion_and_semantic_b200.modules as ours
    import online_joint_depthfusion_and_semantic_b200.modules.extractor as o_ex
    import online_joint_depthfusion_and_semantic_b200.modules.integrator as o_in
    import online_joint_depthfusion_and_semantic_b200.modules.pipeline as o_pl
    sys.modules['modules.pipeline'] = o_pl
    sys.modules['modules.extractor'] = o_ex
    sys.modules['modules.integrator'] = o_in
    return ours


def make_dataset_class(n_frames, h, w, grid, scene_names=('synth0',)):
    """A `Replica`-shaped dataset on the synthetic analytic scene: what utils/setup.get_data + modules/database.py
    consume (`.scenes`, `__len__`, `__getitem__ -> sample`, `get_grid(scene, truncation, semantic_grid)`)."""
    import torch
    from online_joint_depthfusion_and_semantic_b200.synthetic import SyntheticScene

    class SyntheticReplica(torch.utils.data.Dataset):
        def __init__(self, config):
            self.config = config
            self.input = config.input
            self.scenes = list(scene_names)
            self._scene = {s: SyntheticScene(name=s, grid=grid, h=h, w=w, n_frames=n_frames, seed=i, input_key=config.input)
                           for i, s in enumerate(self.scenes)}
            self._index = [(s, i) for s in self.scenes for i in range(n_frames)]

        def __len__(self):
            return len(self._index)

        def __getitem__(self, item):
            s, i = self._index[item]
            b = self._scene[s].frame(i, device='cpu')
            sample = {k: (v[0] if torch.is_tensor(v) else v) for k, v in b.items()}
            sample['frame_id'] = b['frame_id'][0]
            return sample

        def get_grid(self, scene, truncation=None, semantic_grid=False):
            sc = self._scene[scene]
            sdf, lab = sc.gt_volumes(truncation=truncation if truncation is not None else 0.1)
            g = Voxelgrid(sc.resolution)
            g.from_array(sdf.numpy(), sc.bbox)
            if not semantic_grid:
                return (g,)
            l = Voxelgrid(sc.resolution)
            l.from_array(lab.numpy(), sc.bbox)
            return g, l

    return SyntheticReplica


def make_config(workdir, h, w, gpu, strategy='gt', use_semantics=True, n_classes=30, filter_val=2.0):
    """configs/fusion/replica_accuracy.yaml with the synthetic dataset and a test-sized frame."""
    return EasyDict({
        'SETTINGS': {'gpu': bool(gpu), 'num_workers': 0, 'experiment_path': os.path.join(workdir, 'exp'), 'save_mode': 'none',
                     'eval_freq': 2000, 'log_freq': 250, 'seed': 1911, 'implementation': 'efficient'},
        'FUSION_MODEL': {'name': 'v3', 'output_scale': 1.0, 'n_points': 9, 'n_tail_points': 7, 'growth_factor': 6,
                         'use_semantics': bool(use_semantics), 'pretrained': None},
        'SEMANTIC_2D_MODEL': {'stage': 2, 'n_classes': n_classes},
        'TRAINING': {'train_batch_size': 1, 'train_shuffle': False, 'train_ratio': 1, 'val_batch_size': 1, 'val_shuffle': False,
                     'val_ratio': 1},
        'TESTING': {'test_batch_size': 1, 'test_shuffle': False, 'test_ratio': 1, 'outlier_filter_val': filter_val,
                    'fusion_model_path': os.path.join(workdir, 'exp', 'v3', 'model', 'best.pth.tar'),
                    'semantic_2d_model_path': os.path.join(workdir, 'exp', 'adapnet', 'model', 'best.pth.tar')},
        'DATA': {'dataset': 'Replica', 'root_dir': None, 'semantics': 'class30', 'semantic_strategy': strategy, 'semantic_grid': True,
                 'data_load_strategy': None, 'load_scenes_at_once': 1, 'intensity_grad': False, 'input': 'tof_depth',
                 'target': 'depth_gt', 'resx': w, 'resy': h, 'train_scene_list': None, 'val_scene_list': None,
                 'test_scene_list': None, 'init_value': 0.1, 'truncation_strategy': 'standard', 'normalize': True, 'pad': 0},
    })


def write_checkpoints(config, seed=1911):
    """Seeded random-init weights of the REFERENCE's own network classes at the paths the driver loads
    (test_fusion.py:63-71): no trained checkpoint exists offline."""
    import torch
    import modules.model as ref_model
    fm = EasyDict(dict(config.FUSION_MODEL))
    fm.resx, fm.resy = config.DATA.resx, config.DATA.resy
    torch.manual_seed(seed)
    net = ref_model.FusionNet_v3(fm)
    for m in net.modules():                            # non-trivial BatchNorm statistics, like a trained model has
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.05)
            m.running_var.uniform_(0.5, 1.5)
    os.makedirs(os.path.dirname(config.TESTING.fusion_model_path), exist_ok=True)
    torch.save({'model_state': {'_fusion_network.' + k: v for k, v in net.state_dict().items()}}, config.TESTING.fusion_model_path)
    if config.DATA.semantic_strategy == 'predict':
        import modules.adapnet as ref_adapnet
        torch.manual_seed(seed + 1)
        seg = ref_adapnet.AdapNet(config.SEMANTIC_2D_MODEL)
        os.makedirs(os.path.dirname(config.TESTING.semantic_2d_model_path), exist_ok=True)
        torch.save({'model_state': {'module.' + k: v for k, v in seg.state_dict().items()}}, config.TESTING.semantic_2d_model_path)


def drive(impl='reference', gpu=False, frames=12, h=48, w=64, grid=48, strategy='gt', threads=None, workdir=None):
    """Run the reference's own `test_fusion.test_fusion(config)` and return what it logged.

    Returns {'eval': {...}, 'semantic_eval': {...}, 'seconds': wall time of the frame loop + evaluation,
             'reached': 'end' | 'ojdf_call', 'error': str | None}."""
    import logging
    import torch
    enter()
    ours = swap_in_ours() if impl == 'ours' else None
    if threads:
        torch.set_num_threads(int(threads))
    import utils.setup as ref_setup
    import test_fusion as ref_driver                   # the unmodified driver
    ref_setup.Replica = make_dataset_class(frames, h, w, grid)
    workdir = workdir or tempfile.mkdtemp(prefix='ojdf_drive_')
    config = make_config(workdir, h, w, gpu, strategy=strategy)
    write_checkpoints(config)
    if impl == 'ours':
        assert ref_driver.Pipeline is ours.pipeline.Pipeline, 'option A swap did not take'
    if strategy == 'predict':                          # deterministic runs: the eval-time-active dropout off
        orig_pipeline = ref_driver.Pipeline

        def _no_dropout(cfg):
            p = orig_pipeline(cfg)
            for m in p._semantic_2d_network.modules():
                if hasattr(m, 'dropout') and isinstance(m.dropout, bool):
                    m.dropout = False
            return p
        ref_driver.Pipeline = _no_dropout
    captured = {'eval': {}, 'semantic_eval': {}}

    class _Grab(logging.Handler):
        section = None

        def emit(self, record):
            msg = record.getMessage()
            if msg.startswith('Average test results'):
                self.section = 'eval'
            elif msg.startswith('Average semantic results'):
                self.section = 'semantic_eval'
            elif msg.startswith('Per scene'):
                self.section = None
            elif self.section and ':' in msg:
                k, v = msg.split(':', 1)
                try:
                    captured[self.section][k.strip()] = float(v)
                except ValueError:
                    pass

    orig_get_logger = ref_setup.get_logger

    def get_logger(path, name='training'):
        lg = orig_get_logger(path, name)
        lg.addHandler(_Grab())
        return lg
    ref_setup.get_logger = get_logger
    t0 = time.perf_counter()
    reached, err = 'end', None
    try:
        ref_driver.test_fusion(config)
    except Exception as e:                              # noqa: BLE001
        from online_joint_depthfusion_and_semantic_b200._lib import OjdfError
        if impl == 'ours' and isinstance(e, OjdfError):
            reached, err = 'ojdf_call', str(e)          # no GPU here: the drive stops exactly at the CUDA boundary
        else:
            raise
    captured.update(seconds=time.perf_counter() - t0, reached=reached, error=err, impl=impl, frames=frames, h=h, w=w, grid=grid,
                    strategy=strategy, reference=ref_root())
    return captured


def reference_fps(frames=3, warmup=1, h=240, w=320, grid=256, threads=None):
    """Frames per second of the reference's own Pipeline.fuse (modules/pipeline.py:173-248: AdapNet++ stage 2 +
    Extractor + FusionNet_v3 + Integrator) on the host CPU -- the reference arm of bench.py."""
    import torch
    enter()
    import modules.pipeline as ref_pipeline
    from online_joint_depthfusion_and_semantic_b200.synthetic import SyntheticScene
    if threads:
        torch.set_num_threads(int(threads))
    cfg = make_config(tempfile.gettempdir(), h, w, gpu=False, strategy='predict')
    cfg.SETTINGS.device = torch.device('cpu')
    torch.manual_seed(1911)
    pipe = ref_pipeline.Pipeline(cfg).eval()
    scene = SyntheticScene(grid=grid, h=h, w=w, n_frames=max(frames + warmup, 8), intrinsics='replica')
    sdf, lab = scene.gt_volumes()

    class DB:
        pass
    db = DB()
    G = grid
    db.state, vg = {}, lambda a: types.SimpleNamespace(volume=a)
    db.scenes_est = {scene.name: vg(torch.full((G, G, G), 0.1, dtype=torch.float16))}
    db.fusion_weights = {scene.name: torch.zeros((G, G, G), dtype=torch.float16)}
    db.ids_est = {scene.name: vg(torch.zeros((G, G, G), dtype=torch.uint8))}
    db.scores = {scene.name: vg(torch.zeros((G, G, G), dtype=torch.float16))}
    origin = torch.from_numpy(scene.origin)
    db.__class__.__getitem__ = lambda self, s: dict(origin=origin, resolution=scene.resolution, gt=sdf,
                                                    current=self.scenes_est[s].volume, weights=self.fusion_weights[s],
                                                    ids_est=self.ids_est[s].volume, scores=self.scores[s].volume)
    times = []
    with torch.no_grad():
        for i in range(frames + warmup):
            b = scene.frame(i, device='cpu')
            t0 = time.perf_counter()
            pipe.fuse(b, db, torch.device('cpu'))
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return dict(fps=len(times) / sum(times), s_per_frame=sum(times) / len(times), frames=len(times), threads=torch.get_num_threads(),
                reference=ref_root())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('cmd', choices=['drive', 'fps'])
    ap.add_argument('--impl', default='reference', choices=['reference', 'ours'])
    ap.add_argument('--gpu', type=int, default=0)
    ap.add_argument('--frames', type=int, default=12)
    ap.add_argument('--h', type=int, default=48)
    ap.add_argument('--w', type=int, default=64)
    ap.add_argument('--grid', type=int, default=48)
    ap.add_argument('--strategy', default='gt')
    ap.add_argument('--threads', type=int, default=0)
    ap.add_argument('--out', default='')
    a = ap.parse_args()
    if a.cmd == 'drive':
        res = drive(a.impl, bool(a.gpu), a.frames, a.h, a.w, a.grid, a.strategy, a.threads or None)
    else:
        res = reference_fps(a.frames, 1, a.h, a.w, a.grid, a.threads or None)
    txt = json.dumps(res)
    if a.out:
        with open(a.out, 'w') as f:
            f.write(txt)
    print('HARNESS_RESULT ' + txt)


if __name__ == '__main__':
    main()
