"""Database -- the per-scene voxel state the hot path reads and writes (reference:
modules/database.py:18-103,108-116,351-421).

Volume contract (a16): per scene `scenes_est[s].volume` f16 (X,Y,Z) initialised to +init_value,
`fusion_weights[s]` f16 zeros, `ids_est[s].volume` u8 zeros, `scores[s].volume` f16 zeros,
`scenes_gt[s].volume` f16, `ids_gt[s].volume` u8, `origin[s]` f64 (3,) tensor, `resolution[s]` float;
dense row-major, z fastest.  Volumes live on the GPU for their whole life (the reference's
`efficient` implementation, modules/database.py:408-421); the mesh / HDF5 / PLY writers
(modules/database.py:118-261) are out of scope.

`dataset` only needs `.scenes` and `get_grid(scene, truncation, semantic_grid) -> (grid[, labels])`
with `.volume`, `.origin`, `.resolution`, `.bbox` -- what modules/database.py:48-76 consumes.
"""
import numpy as np
import torch

from . import metrics


def _np(x):
    """numpy view of a tensor (device -> host copy) or of an array."""
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


class Voxelgrid:
    """The four attributes of deps/graphics Voxelgrid the hot path touches (voxelgrid.py:54-70,157-161)."""

    def __init__(self, resolution):
        self.resolution = resolution
        self.volume = None
        self.bbox = None
        self.origin = None

    def from_array(self, array, bbox):
        self.volume = array
        self.bbox = np.asarray(bbox)
        self.origin = self.bbox[:, 0].copy()

    @property
    def shape(self):
        return tuple(self.volume.shape)


class Database:

    def __init__(self, dataset, config):
        self.device = config.device
        self.implementation = getattr(config, 'implementation', 'efficient')
        self.initial_value = config.init_value
        self.semantics = config.semantics
        self.semantic_grid = getattr(config, 'semantic_grid', False)
        if self.semantics:
            self.n_classes = getattr(config, 'n_classes', None)
        self.scenes, self.state, self.origin, self.resolution = [], {}, {}, {}
        self.scenes_gt, self.scenes_est, self.fusion_weights = {}, {}, {}
        self.ids_gt, self.ids_est, self.scores = {}, {}, {}
        for s in dataset.scenes:
            grid = dataset.get_grid(s, self.initial_value, self.semantic_grid)
            self.scenes.append(s)
            self.scenes_gt[s] = grid[0]
            self.scenes_gt[s].volume = self._dev(grid[0].volume, torch.float16)
            self.origin[s] = torch.as_tensor(np.asarray(grid[0].origin, dtype=np.float64))
            self.resolution[s] = float(grid[0].resolution)
            self.scenes_est[s] = Voxelgrid(self.resolution[s])
            self.scenes_est[s].bbox, self.scenes_est[s].origin = grid[0].bbox, grid[0].origin
            if self.semantics:
                if self.semantic_grid:
                    self.ids_gt[s] = grid[1]
                    self.ids_gt[s].volume = self._dev(grid[1].volume, torch.uint8)
                self.ids_est[s] = Voxelgrid(self.resolution[s])
                self.scores[s] = Voxelgrid(self.resolution[s])
        self.reset()

    def _dev(self, a, dtype):
        t = torch.as_tensor(a) if not torch.is_tensor(a) else a
        return t.to(device=self.device, dtype=dtype).contiguous()

    def __getitem__(self, item):
        sample = dict(origin=self.origin[item], resolution=self.resolution[item], gt=self.scenes_gt[item].volume,
                      current=self.scenes_est[item].volume, weights=self.fusion_weights[item])
        if self.semantics:
            sample['ids_est'] = self.ids_est[item].volume
            sample['scores'] = self.scores[item].volume
            if self.semantic_grid:
                sample['ids_gt'] = self.ids_gt[item].volume
        else:
            sample.update(histograms=None, ids_est=None, ids_gt=None, scores=None)
        return sample

    def __len__(self):
        return len(self.scenes_gt)

    def reset(self, scene_id=None):
        """modules/database.py:351-371, without the numpy round trip."""
        for s in ([scene_id] if scene_id else self.scenes):
            shape = tuple(self.scenes_gt[s].volume.shape)
            self.state[s] = False
            self.scenes_est[s].volume = torch.full(shape, self.initial_value, dtype=torch.float16, device=self.device)
            self.fusion_weights[s] = torch.zeros(shape, dtype=torch.float16, device=self.device)
            if self.semantics:
                self.ids_est[s].volume = torch.zeros(shape, dtype=torch.uint8, device=self.device)
                self.scores[s].volume = torch.zeros(shape, dtype=torch.float16, device=self.device)

    # ---- host <-> device round trip (modules/database.py:383-421).  The reference's drivers call to_numpy() before the
    # filters / metrics (test_fusion.py:82, train_fusion.py:198,215) and to_torch() before fusing again
    # (train_fusion.py:254).  Here the filters and metrics run on the device either way: they accept whatever form the
    # volumes are in and leave them in that form.
    def to_numpy(self):
        for s in self.scenes:
            self.origin[s] = _np(self.origin[s])
            self.scenes_est[s].volume = _np(self.scenes_est[s].volume)
            self.scenes_gt[s].volume = _np(self.scenes_gt[s].volume)
            self.fusion_weights[s] = _np(self.fusion_weights[s])
            if self.semantics:
                self.ids_est[s].volume = _np(self.ids_est[s].volume)
                self.scores[s].volume = _np(self.scores[s].volume)
                if self.semantic_grid:
                    self.ids_gt[s].volume = _np(self.ids_gt[s].volume)

    def to_torch(self, gt=True, scenes=None):
        """Back to device tensors (every volume lives on `self.device`: the `efficient` implementation,
        modules/database.py:408-421; the reference's `standard` mode, which re-uploads both volumes every frame, is not
        offered).  gt=False leaves origin / GT volumes as they are, like the reference."""
        for s in (self.scenes if scenes is None else [scenes]):
            self.scenes_est[s].volume = self._dev(self.scenes_est[s].volume, torch.float16)
            self.fusion_weights[s] = self._dev(self.fusion_weights[s], torch.float16)
            if gt:
                self.origin[s] = torch.as_tensor(np.asarray(_np(self.origin[s]), dtype=np.float64))
                self.scenes_gt[s].volume = self._dev(self.scenes_gt[s].volume, torch.float16)
            if self.semantics:
                self.ids_est[s].volume = self._dev(self.ids_est[s].volume, torch.uint8)
                self.scores[s].volume = self._dev(self.scores[s].volume, torch.float16)
                if gt and self.semantic_grid:
                    self.ids_gt[s].volume = self._dev(self.ids_gt[s].volume, torch.uint8)

    def _like(self, result, original):
        """`result` (device tensor) in the form `original` was stored in."""
        return _np(result) if isinstance(original, np.ndarray) else result

    def filter(self, value=2.):
        """modules/database.py:108-112."""
        for s in self.scenes:
            est0, w0 = self.scenes_est[s].volume, self.fusion_weights[s]
            est, w = self._dev(est0, torch.float16), self._dev(w0, torch.float16)
            low = w < value
            est[low] = self.initial_value
            w[low] = 0
            if est is not est0:                              # host-resident (after to_numpy()): store back in that form
                self.scenes_est[s].volume = self._like(est, est0)
            if w is not w0:
                self.fusion_weights[s] = self._like(w, w0)

    def filter_semantics(self, value=5):
        """modules/database.py:114-116 (scipy median filter of the label volume), on the device."""
        for s in self.scenes:
            ids0 = self.ids_est[s].volume
            self.ids_est[s].volume = self._like(metrics.median_filter_labels(self._dev(ids0, torch.uint8), size=value), ids0)

    def evaluate(self, mode='train', workspace=None):
        """modules/database.py:265-310: mse / mad / iou / acc (+ 'f1') averaged over the scenes, without leaving
        the device.  mode == 'test' also returns the per-scene results."""
        total, per_scene = {}, {}
        for s in self.scenes:
            if not self.state[s]:
                continue
            r = metrics.evaluation(self._dev(self.scenes_est[s].volume, torch.float16),
                                   self._dev(self.scenes_gt[s].volume, torch.float16),
                                   self._dev(self.fusion_weights[s], torch.float16) > 0)
            per_scene[s] = r
            for k, v in r.items():
                if workspace is not None:
                    workspace.log('{} {}'.format(k, v), mode)
                total[k] = total.get(k, 0.0) + v
        for k in total:
            total[k] /= len(self.scenes_est)
        return (total, per_scene) if mode == 'test' else total

    def evaluate_semantics(self, mode='train', workspace=None):
        """modules/database.py:312-349: mean class accuracy / mean IoU of the label volumes."""
        total, per_scene = {}, {}
        for s in self.scenes:
            if not self.state[s]:
                continue
            r, cls_iou = metrics.semantic_evaluation(self._dev(self.ids_est[s].volume, torch.uint8),
                                                     self._dev(self.ids_gt[s].volume, torch.uint8),
                                                     self._dev(self.fusion_weights[s], torch.float16) > 0, self.n_classes)
            per_scene[s] = cls_iou
            for k, v in r.items():
                if workspace is not None:
                    workspace.log('{} {}'.format(k, v), mode)
                total[k] = total.get(k, 0.0) + v
        for k in total:
            total[k] /= len(self.scenes_est)
        return total, per_scene

    def save(self, path=None, save_mode=None, scene_id=None):
        """modules/database.py:180-261 writes HDF5 / PLY files with h5py, skimage and trimesh -- out of scope here (I/O,
        SURVEY.md section 2 row 6).  Kept as an explicit no-op for the reference's 'test' call sites with an unknown mode
        (the reference itself does nothing for modes other than ply / tsdf / test); anything else raises."""
        if save_mode in ('ply', 'tsdf', 'test'):
            raise NotImplementedError("Database.save(%r): mesh / HDF5 export is out of scope of the B200 fusion path; "
                                      "use the reference's writers on database.to_numpy() volumes" % (save_mode,))
