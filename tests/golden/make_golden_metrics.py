"""Golden fixture for modules/metrics.py, produced by the REFERENCE's own utils/metrics.py:69-196 and by
scipy.ndimage.median_filter as modules/database.py:114-116 calls it (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_metrics.py

Stores seeded input volumes (fp16 TSDF with NaNs, fp16 weights, u8 labels) and the reference's outputs."""
import json
import os
import sys

import numpy as np
from scipy.ndimage import median_filter

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    mg.ref_env()
    import utils.metrics as ref_metrics
    out = {}
    arrays = {}
    for case, (seed, shape, n_class) in enumerate(((3, (24, 19, 31), 30), (4, (16, 16, 16), 21))):
        rng = np.random.default_rng(seed)
        est = (0.06 * rng.standard_normal(shape)).astype(np.float16)
        gt = (0.06 * rng.standard_normal(shape)).astype(np.float16)
        est[rng.random(shape) < 0.01] = np.nan
        w = ((rng.random(shape) < 0.6) * rng.random(shape) * 8).astype(np.float16)
        ids_gt = rng.integers(0, n_class // 2, shape).astype(np.uint8)
        ids_est = np.where(rng.random(shape) < 0.7, ids_gt, rng.integers(0, n_class, shape)).astype(np.uint8)
        mask = w > 0
        ev = ref_metrics.evaluation(est, gt, mask)
        sem, cls_iou = ref_metrics.semantic_evaluation(ids_est, ids_gt, mask, n_class)
        med = median_filter(ids_est, size=5)
        arrays.update({'c%d_est' % case: est.view(np.uint16), 'c%d_gt' % case: gt.view(np.uint16), 'c%d_w' % case: w.view(np.uint16),
                       'c%d_ids_est' % case: ids_est, 'c%d_ids_gt' % case: ids_gt, 'c%d_median5' % case: med})
        out['case%d' % case] = dict(n_class=n_class, evaluation={k: float(v) for k, v in ev.items()},
                                    semantic={k: float(v) for k, v in sem.items()},
                                    class_iou={str(int(k)): float(v) for k, v in cls_iou.items()})
    np.savez_compressed(os.path.join(HERE, 'metrics_volumes.npz'), **arrays)
    json.dump(out, open(os.path.join(HERE, 'metrics.json'), 'w'), indent=1)
    print(json.dumps(out)[:400])


if __name__ == '__main__':
    main()
