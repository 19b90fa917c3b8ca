"""Boundary check (INTEGRATION.md option A): the UNMODIFIED reference driver `test_fusion.py` runs in
this image through baseline/harness.py (stand-ins for easydict / h5py / trimesh / skimage / graphics, a synthetic
`Replica` dataset injected into utils.setup), (a) with the reference's own modules end to end on the CPU and (b) with
this repository's `modules.pipeline / extractor / integrator` swapped in -- which, without a GPU, must get exactly as
far as the first CUDA call and fail loudly there (no CPU fallback).  Needs the reference tree: /root/reference in
the build container, baseline/_ref on the GPU box (python baseline/install_ref.py)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import harness  # noqa: E402

needs_ref = pytest.mark.skipif(harness.ref_root() is None, reason='reference tree not available (baseline/_ref missing)')


def run_harness(*args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'baseline', 'harness.py')] + list(args), capture_output=True,
                       text=True, timeout=timeout, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith('HARNESS_RESULT ')]
    assert r.returncode == 0 and lines, (r.stdout[-2000:], r.stderr[-2000:])
    return json.loads(lines[-1][len('HARNESS_RESULT '):])


@needs_ref
@pytest.mark.timeout(600)
def test_unmodified_reference_driver_runs_end_to_end_on_cpu():
    res = run_harness('drive', '--impl', 'reference', '--gpu', '0', '--frames', '4')
    assert res['reached'] == 'end'
    assert set(res['eval']) >= {'mse', 'mad', 'iou', 'acc'}                 # utils/metrics.py:111-196 via Database.evaluate
    assert set(res['semantic_eval']) >= {'Mean IoU', 'Mean Acc'}            # utils/metrics.py:69-108
    assert 0.0 <= res['eval']['iou'] <= 1.0 and res['eval']['mse'] > 0.0


@needs_ref
@pytest.mark.timeout(600)
def test_option_a_swap_reaches_the_cuda_boundary_and_fails_loudly_on_cpu():
    """With modules.pipeline/extractor/integrator swapped for ours, the reference's driver builds its dataset, its
    Database, OUR Pipeline, loads the reference-format checkpoint into it (test_fusion.py:63-65) and enters the frame
    loop; on CPU tensors the first libojdf call refuses (the product path has no fallback)."""
    res = run_harness('drive', '--impl', 'ours', '--gpu', '0', '--frames', '4')
    assert res['reached'] == 'ojdf_call'
    assert 'no CPU fallback' in res['error']
