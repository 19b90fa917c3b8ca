"""modules/metrics.py (torch, device-agnostic) against a numpy restatement of the reference's utils/metrics.py:69-196
and against scipy's median filter (modules/database.py:114-116)."""
import numpy as np
import pytest
import torch
from scipy.ndimage import median_filter

from online_joint_depthfusion_and_semantic_b200.modules import metrics


def _np_evaluation(est, target, mask):
    """utils/metrics.py:110-196, line by line."""
    eps = 1.e-10
    est = np.clip(np.nan_to_num(est.astype(np.float32)), -0.04, 0.04)
    target = np.clip(np.nan_to_num(target.astype(np.float32)), -0.04, 0.04)
    mse = np.nansum(mask * np.power(est - target, 2)) / (np.nansum(mask) + eps)
    mad = np.nansum((mask * np.abs(est - target)).astype(np.float32)) / (np.nansum(mask) + eps)
    tp = (est < 0) & (target < 0) & (mask > 0)
    fp = (est < 0) & (target >= 0) & (mask > 0)
    fn = (est >= 0) & (target < 0) & (mask > 0)
    tn = (est >= 0) & (target >= 0) & (mask > 0)
    iou = np.nansum(tp) / (np.nansum(tp) + np.nansum(fp) + np.nansum(fn) + eps)
    acc = (np.nansum(tp) + np.nansum(tn)) / (np.nansum(mask) + eps)
    return {'mse': mse, 'mad': mad, 'iou': iou, 'acc': acc}


def _np_semantic(est, target, mask, n_class):
    """utils/metrics.py:69-108, line by line."""
    eps = np.finfo(np.float32).eps
    est = est.flatten() * mask.flatten()
    target = target.flatten() * mask.flatten()
    est_ids = np.bincount(np.unique(est.flatten()), minlength=n_class)
    gt_ids = np.bincount(np.unique(target.flatten()), minlength=n_class)
    m = (target >= 0) & (target < n_class)
    hist = np.bincount(n_class * target[m].astype(np.uint16) + est[m], minlength=n_class * n_class).reshape(n_class, n_class)
    tp = np.diag(hist)
    fp = hist.sum(axis=0) - tp
    fn = hist.sum(axis=1) - tp
    valid_ids = np.sum(gt_ids) - 1
    acc = tp / (tp + fn + eps)
    iou = tp / (tp + fn + fp + eps)
    valid = np.where(est_ids | gt_ids)[0]
    return {'Mean Acc': np.sum(acc[1:]) / valid_ids, 'Mean IoU': np.sum(iou[1:]) / valid_ids}, dict(zip(valid, iou[valid]))


@pytest.mark.parametrize('seed', [0, 1])
def test_evaluation_matches_reference_formulas(seed):
    rng = np.random.default_rng(seed)
    shape = (24, 19, 31)
    est = (0.06 * rng.standard_normal(shape)).astype(np.float16)
    gt = (0.06 * rng.standard_normal(shape)).astype(np.float16)
    est[rng.random(shape) < 0.01] = np.nan
    w = (rng.random(shape) < 0.6) * rng.random(shape)
    mask = w > 0
    ref = _np_evaluation(est, gt, mask)
    got = metrics.evaluation(torch.from_numpy(est), torch.from_numpy(gt), torch.from_numpy(mask))
    for k in ref:
        assert abs(got[k] - ref[k]) <= 2e-6 * max(abs(ref[k]), 1e-12), (k, got[k], ref[k])
    tp = float(((np.nan_to_num(est.astype(np.float32)) < 0) & (gt < 0) & mask).sum())
    assert 0.0 < got['f1'] <= 1.0 and got['f1'] >= got['iou'] and tp > 0


@pytest.mark.parametrize('n_class', [8, 30])
def test_semantic_evaluation_matches_reference_formulas(n_class):
    rng = np.random.default_rng(n_class)
    shape = (20, 22, 18)
    gt = rng.integers(0, n_class - 2, shape).astype(np.uint8)          # the two highest labels never occur
    est = np.where(rng.random(shape) < 0.7, gt, rng.integers(0, n_class, shape)).astype(np.uint8)
    mask = rng.random(shape) < 0.5
    ref, ref_cls = _np_semantic(est, gt, mask, n_class)
    got, got_cls = metrics.semantic_evaluation(torch.from_numpy(est), torch.from_numpy(gt), torch.from_numpy(mask), n_class)
    for k in ref:
        assert abs(got[k] - ref[k]) <= 1e-9, (k, got[k], ref[k])
    assert sorted(got_cls) == sorted(int(c) for c in ref_cls)
    for c in got_cls:
        assert abs(got_cls[c] - ref_cls[c]) <= 1e-9


@pytest.mark.parametrize('shape,size', [((17, 12, 21), 5), ((9, 9, 9), 3), ((6, 30, 7), 5)])
def test_median_filter_is_bit_identical_to_scipy(shape, size):
    rng = np.random.default_rng(sum(shape))
    ids = (rng.integers(0, 12, shape) * (rng.random(shape) < 0.7)).astype(np.uint8)
    ref = median_filter(ids, size=size)
    got = metrics.median_filter_labels(torch.from_numpy(ids), size=size).numpy()
    assert got.dtype == np.uint8 and np.array_equal(got, ref)


class _Grid:
    def __init__(self, vol, res=0.05):
        self.volume, self.resolution = vol, res
        self.bbox = np.array([[0.0, 1.0]] * 3)
        self.origin = self.bbox[:, 0].copy()


class _Dataset:
    scenes = ['a', 'b']

    def __init__(self, rng, shape):
        self.g = {s: (_Grid(np.clip(0.08 * rng.standard_normal(shape), -0.1, 0.1).astype(np.float16)),
                      _Grid(rng.integers(0, 6, shape).astype(np.uint8))) for s in self.scenes}

    def get_grid(self, s, truncation, semantic_grid):
        return self.g[s]


def test_database_evaluate_protocol_on_cpu():
    """Database.evaluate / evaluate_semantics / filter_semantics (modules/database.py:108-116,265-349) on device tensors:
    same keys, same averaging over all scenes, only scenes with integrated frames contribute."""
    from types import SimpleNamespace
    from online_joint_depthfusion_and_semantic_b200.modules.database import Database
    rng = np.random.default_rng(4)
    shape = (12, 10, 14)
    ds = _Dataset(rng, shape)
    db = Database(ds, SimpleNamespace(device='cpu', implementation='efficient', init_value=0.1, semantics='class30',
                                      semantic_grid=True, n_classes=6))
    assert db.evaluate() == {}                                   # nothing integrated yet
    for s in db.scenes:
        db.scenes_est[s].volume = torch.from_numpy((0.08 * rng.standard_normal(shape)).astype(np.float16))
        db.fusion_weights[s] = torch.from_numpy(((rng.random(shape) < 0.5) * 3.0).astype(np.float16))
        db.ids_est[s].volume = torch.from_numpy(rng.integers(0, 6, shape).astype(np.uint8))
    db.state['a'] = True
    one = db.evaluate()
    ref = _np_evaluation(db.scenes_est['a'].volume.numpy(), db.scenes_gt['a'].volume.numpy(), (db.fusion_weights['a'] > 0).numpy())
    assert set(one) == {'mse', 'mad', 'iou', 'acc', 'f1'}
    assert abs(one['iou'] - ref['iou'] / 2) < 1e-9              # divided by the number of scenes, as the reference does
    db.state['b'] = True
    total, per_scene = db.evaluate(mode='test')
    assert set(per_scene) == {'a', 'b'} and abs(total['mse'] - (per_scene['a']['mse'] + per_scene['b']['mse']) / 2) < 1e-12
    sem, cls = db.evaluate_semantics(mode='test')
    assert set(sem) == {'Mean Acc', 'Mean IoU'} and set(cls) == {'a', 'b'}
    before = db.ids_est['a'].volume.clone()
    db.filter_semantics(5)
    assert np.array_equal(db.ids_est['a'].volume.numpy(), median_filter(before.numpy(), size=5))


def _check_against_reference_golden(device):
    """modules/metrics.py vs the REFERENCE's own utils/metrics.py + scipy median filter (tests/golden/make_golden_metrics.py)."""
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    gold = json.load(open(os.path.join(here, 'metrics.json')))
    vol = np.load(os.path.join(here, 'metrics_volumes.npz'))
    for case in range(2):
        g = gold['case%d' % case]
        t = lambda k, dt: torch.from_numpy(vol['c%d_%s' % (case, k)].copy()).view(dt).to(device)   # noqa: E731
        est, gt, w = t('est', torch.float16), t('gt', torch.float16), t('w', torch.float16)
        ids_est, ids_gt = t('ids_est', torch.uint8), t('ids_gt', torch.uint8)
        ev = metrics.evaluation(est, gt, w > 0)
        for k, v in g['evaluation'].items():
            assert abs(ev[k] - v) <= 2e-6 * max(abs(v), 1e-12), (case, k, ev[k], v)
        sem, cls_iou = metrics.semantic_evaluation(ids_est, ids_gt, w > 0, g['n_class'])
        for k, v in g['semantic'].items():
            assert abs(sem[k] - v) <= 1e-6 * max(abs(v), 1e-12), (case, k, sem[k], v)
        assert {str(int(k)) for k in cls_iou} == set(g['class_iou'])
        for k, v in cls_iou.items():
            assert abs(float(v) - g['class_iou'][str(int(k))]) <= 1e-6
        med = metrics.median_filter_labels(ids_est, size=5)
        assert np.array_equal(med.cpu().numpy(), vol['c%d_median5' % case])


def test_metrics_vs_reference_golden_cpu():
    _check_against_reference_golden(torch.device('cpu'))


@pytest.mark.gpu
def test_metrics_vs_reference_golden_gpu():
    """Row f1 on the device the volumes live on."""
    _check_against_reference_golden(torch.device('cuda:0'))
