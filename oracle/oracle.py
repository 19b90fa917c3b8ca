"""ctypes binding of oracle/ojdf_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module (as the checker / reported CPU baseline).  The product
package never imports it; the product path is the CUDA library and raises without it.

Arrays are numpy; fp16 volumes travel as uint16 bit patterns (np.float16 views work).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libojdf_oracle.so')
_lib = None


def build(force=False):
    src = os.path.join(_HERE, 'ojdf_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-B', 'libojdf_oracle.so'], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.ojdf_oracle_integrate.restype = C.c_int
        _lib.ojdf_oracle_integrate_frame.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    a = np.ascontiguousarray(a, dtype=dt)
    return a


def _u16(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float16:
        a = a.view(np.uint16)
    assert a.dtype == np.uint16
    return a


def set_threads(n):
    lib().ojdf_oracle_set_threads(int(n))


def max_threads():
    return int(lib().ojdf_oracle_max_threads())


def unproject(depth, Kinv, E, fma_chain=True):
    """depth (h,w) f32, Kinv (3,3) f32, E (3|4,4) f32 -> world (h*w,3) f32."""
    depth = _c(depth, np.float32)
    h, w = depth.shape
    Kinv = _c(Kinv, np.float32)
    E = _c(np.asarray(E)[:3], np.float32)
    world = np.empty((h * w, 3), np.float32)
    lib().ojdf_oracle_unproject(_p(depth), h, w, _p(Kinv), _p(E), int(bool(fma_chain)), _p(world))
    return world


def extract(world, eye, origin, res, tsdf, wvol, P=9, full=False):
    """world (N,3) f32 -> dict(fusion_values, fusion_weights[, points, indices, weights])."""
    world = _c(world, np.float32)
    N = world.shape[0]
    eye = _c(eye, np.float32)
    origin = _c(origin, np.float64)
    tsdf, wvol = _u16(tsdf), _u16(wvol)
    X, Y, Z = tsdf.shape
    vals = np.empty((N, P), np.float32)
    wts = np.empty((N, P), np.float32)
    pts = np.empty((N, P, 3), np.float64) if full else None
    idx = np.empty((N, P, 8, 3), np.int64) if full else None
    cw = np.empty((N, P, 8), np.float64) if full else None
    lib().ojdf_oracle_extract(_p(world), C.c_int64(N), _p(eye), _p(origin), C.c_double(float(res)),
                              _p(tsdf), _p(wvol), X, Y, Z, P, _p(vals), _p(wts), _p(pts), _p(idx), _p(cw))
    out = dict(fusion_values=vals, fusion_weights=wts)
    if full:
        out.update(points=pts, indices=idx, weights=cw)
    return out


def integrate(values, indices, weights, tsdf, wvol, ids=None, scores=None, ids_vol=None, scores_vol=None,
              do_sem=False):
    """Reference `updates` form; volumes (uint16 / uint8 arrays) are updated IN PLACE."""
    values = _c(values, np.float32).reshape(-1)
    M1 = values.shape[0]
    indices = _c(indices, np.int64).reshape(M1, 8, 3)
    weights = _c(weights, np.float64).reshape(M1, 8)
    assert tsdf.dtype in (np.uint16, np.float16) and tsdf.flags.c_contiguous and wvol.flags.c_contiguous
    X, Y, Z = tsdf.shape
    if do_sem:
        ids = _c(ids, np.uint8).reshape(-1)
        scores = _c(scores, np.float32).reshape(-1)
        assert ids.shape[0] == M1 and scores.shape[0] == M1
        assert ids_vol.dtype == np.uint8 and ids_vol.flags.c_contiguous and scores_vol.flags.c_contiguous
    rc = lib().ojdf_oracle_integrate(_p(values), _p(indices), _p(weights), C.c_int64(M1), _p(tsdf), _p(wvol),
                                     X, Y, Z, _p(ids), _p(scores), _p(ids_vol), _p(scores_vol), int(bool(do_sem)))
    if rc != 0:
        raise MemoryError('ojdf_oracle_integrate rc=%d' % rc)


def integrate_frame(world, filt_depth, est, eye, origin, res, tsdf, wvol, tail=7, clampv=0.1,
                    pix_ids=None, pix_scores=None, ids_vol=None, scores_vol=None, do_sem=False):
    """Whole-frame form (modules/pipeline.py:137-171 + modules/integrator.py:15-126), in place."""
    world = _c(world, np.float32)
    N = world.shape[0]
    filt_depth = _c(filt_depth, np.float32).reshape(-1)
    est = _c(est, np.float32).reshape(N, -1)
    P = est.shape[1]
    eye = _c(eye, np.float32)
    origin = _c(origin, np.float64)
    assert tsdf.flags.c_contiguous and wvol.flags.c_contiguous
    X, Y, Z = tsdf.shape
    if do_sem:
        pix_ids = _c(pix_ids, np.uint8).reshape(-1)
        pix_scores = _c(pix_scores, np.float32).reshape(-1)
    rc = lib().ojdf_oracle_integrate_frame(_p(world), _p(filt_depth), _p(est), C.c_int64(N), _p(eye), _p(origin),
                                           C.c_double(float(res)), P, int(tail), C.c_float(float(np.float32(clampv))),
                                           _p(tsdf), _p(wvol), X, Y, Z, _p(pix_ids), _p(pix_scores),
                                           _p(ids_vol), _p(scores_vol), int(bool(do_sem)))
    if rc != 0:
        raise MemoryError('ojdf_oracle_integrate_frame rc=%d' % rc)
