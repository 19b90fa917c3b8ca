"""On-device restatement of the reference's volume metrics (utils/metrics.py:69-196) and of the label
median filter (modules/database.py:114-116, scipy.ndimage.median_filter) in plain torch ops.

The reference moves every volume to the host (`Database.to_numpy`, train_fusion.py:198,215) and evaluates
with numpy / scipy; here the volumes stay where the integrator left them.  Sums are
taken in float64, so values agree with the float32 numpy originals to ~1e-6 relative, counts exactly.
Also provides the occupancy F1 built from the very same tp / fp / fn masks as `iou_fn`
(utils/metrics.py:164-181), which SURVEY.md 8d names as the F-score of the parity report."""
import torch

_EPS = 1.e-10
_EPS32 = float(torch.finfo(torch.float32).eps)


def _prep(est, target):
    """evaluation() (utils/metrics.py:110-116): float32, NaN -> 0, clip to +-0.04."""
    est = torch.nan_to_num(est.float()).clamp(-0.04, 0.04)
    target = torch.nan_to_num(target.float()).clamp(-0.04, 0.04)
    return est, target


def evaluation(est, target, mask=None):
    """mse / mad / iou / acc of a TSDF volume against the ground truth (utils/metrics.py:110-196), plus 'f1'
    (occupancy F1 from the iou masks).  est / target: (X,Y,Z) tensors on any device; mask: bool or None."""
    est, target = _prep(est, target)
    d = (est - target).double()
    if mask is None:
        n = float(d.numel())
        m = torch.ones_like(est, dtype=torch.bool)
        mse, mad = float((d * d).sum()) / n, float(d.abs().sum()) / n
        denom_acc = n + _EPS
    else:
        m = mask.bool()
        w = m.double()
        cnt = float(w.sum())
        mse = float((w * d * d).sum()) / (cnt + _EPS)
        mad = float((w * d.abs().float().double()).sum()) / (cnt + _EPS)
        denom_acc = cnt + _EPS
    occ_e, occ_t = est < 0, target < 0
    tp = float((occ_e & occ_t & m).sum())
    fp = float((occ_e & ~occ_t & m).sum())
    fn = float((~occ_e & occ_t & m).sum())
    tn = float((~occ_e & ~occ_t & m).sum())
    return {'mse': mse, 'mad': mad, 'iou': tp / (tp + fp + fn + _EPS), 'acc': (tp + tn) / denom_acc,
            'f1': 2.0 * tp / (2.0 * tp + fp + fn + _EPS)}


def semantic_evaluation(est, target, mask, n_class):
    """Mean class accuracy / mean IoU over the labels present in the ground truth, label 0 excluded
    (utils/metrics.py:69-108).  est / target: integer label volumes, mask: bool volume."""
    m = mask.reshape(-1).bool()
    e = est.reshape(-1).long() * m
    t = target.reshape(-1).long() * m
    ok = (t >= 0) & (t < n_class)
    hist = torch.bincount(n_class * t[ok] + e[ok], minlength=n_class * n_class)[:n_class * n_class]
    hist = hist.reshape(n_class, n_class).double()              # target x estimate
    tp = hist.diag()
    fp = hist.sum(0) - tp
    fn = hist.sum(1) - tp
    est_ids = torch.bincount(torch.unique(e), minlength=n_class)[:n_class] > 0
    gt_ids = torch.bincount(torch.unique(t), minlength=n_class)[:n_class] > 0
    valid_ids = float(gt_ids.sum()) - 1.0
    acc = tp / (tp + fn + _EPS32)
    iou = tp / (tp + fn + fp + _EPS32)
    metrics = {'Mean Acc': float(acc[1:].sum()) / valid_ids, 'Mean IoU': float(iou[1:].sum()) / valid_ids}
    valid = torch.nonzero(est_ids | gt_ids).reshape(-1).tolist()
    return metrics, {int(c): float(iou[c]) for c in valid}


def _pad_symmetric(x, p):
    """scipy's mode='reflect' (d c b a | a b c d | d c b a), i.e. numpy 'symmetric', along all three axes."""
    for ax in range(3):
        n = x.shape[ax]
        idx = torch.cat([torch.arange(p - 1, -1, -1), torch.arange(n), torch.arange(n - 1, n - 1 - p, -1)]).clamp(0, n - 1)
        x = x.index_select(ax, idx.to(x.device))
    return x


def median_filter_labels(ids, size=5):
    """scipy.ndimage.median_filter(ids, size) for a uint8 label volume (Database.filter_semantics).
    The median of the size^3 window is the smallest label c with #(window <= c) > size^3 / 2; the counts are
    exact box sums of 0/1 masks, so the result is bit-identical to scipy's."""
    assert size % 2 == 1
    p, half = size // 2, (size ** 3) // 2
    x = _pad_symmetric(ids, p)
    top = int(ids.max()) if ids.numel() else 0
    out = torch.zeros(ids.shape, dtype=torch.int32, device=ids.device)
    for c in range(top):                                        # for c = top the count is the whole window
        le = (x <= c).float()[None, None]
        cnt = torch.nn.functional.avg_pool3d(le, size, stride=1, divisor_override=1)[0, 0]
        out += (cnt <= half).int()
    return out.to(ids.dtype)
