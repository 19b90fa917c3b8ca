"""Extractor -- drop-in for the reference's modules/extractor.py:8-79.

Reference behaviour reproduced (file:line in /root/reference):
  * forward(depth, extrinsics, intrinsics, tsdf_volume, weights_volume, origin, resolution)
    returns a dict with fusion_values / fusion_weights (b,N,P) f32, points (b,N,P,3) f64,
    depth (b,N) f32, indices (b,N,P,8,3) i64, weights (b,N,P,8) f64, pcl (b,N,3) f32
    (modules/extractor.py:24,69-75);
  * intrinsics are cast to f32 and inverted on the host exactly as the reference's CPU path
    does (`intrinsics.float().inverse()`, modules/extractor.py:39,104);
  * batch size is 1 on this path (modules/pipeline.py:199, modules/extractor.py:314-318).

What is different by design: one launch pair (ojdf_extract) instead of ~970 ATen calls, and
the three big tensors (points, indices, weights: 177 MB per 240x320 frame) are only
materialised if somebody actually reads them -- the integrator consumes the compact per-ray
record `ray` (N,6) f64 instead.
"""
import torch
from torch import nn

from .. import _lib

_LAZY = ('points', 'indices', 'weights')


class ExtractedValues(dict):
    """The reference's `values` dict; points/indices/weights appear on first access."""

    def __init__(self, eager, materialise):
        super().__init__(eager)
        self._materialise = materialise

    def __missing__(self, key):
        if key in _LAZY and self._materialise is not None:
            self.update(self._materialise())
            self._materialise = None
            return dict.__getitem__(self, key)
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or (key in _LAZY and self._materialise is not None)


def host_pose(extrinsics, intrinsics):
    """(Kinv 3x3 f32, E 3x4 f32) as contiguous CPU tensors; intrinsics may be None."""
    E = extrinsics.detach().float().cpu().reshape(-1, 4)[:3].contiguous()
    Kinv = None
    if intrinsics is not None:
        Kinv = intrinsics.detach().cpu().float().reshape(3, 3).inverse().float().contiguous()
    return Kinv, E


class Extractor(nn.Module):

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.n_points = config.FUSION_MODEL.n_points
        self.mode = 'ray'

    def forward(self, depth, extrinsics, intrinsics, tsdf_volume, weights_volume, origin, resolution,
                world=None, eager=False):
        """`world` (optional, (b,N,3) f32): use these world points instead of unprojecting
        `depth` (parity tests pass the oracle's `pcl`).  `eager=True` materialises
        points/indices/weights immediately, like the reference always does."""
        b, h, w = depth.shape
        if b != 1:
            raise ValueError('the fusion path is batch-size-1 (modules/pipeline.py:199); got b=%d' % b)
        _lib.require_cuda(depth, tsdf_volume, weights_volume, world)
        if tsdf_volume.dtype != torch.float16 or weights_volume.dtype != torch.float16:
            raise TypeError('volumes must be float16 (modules/database.py:60,64)')
        dev = depth.device
        depth_f = depth.detach().float().contiguous()
        N, P = h * w, int(self.n_points)
        Kinv, E = host_pose(extrinsics, intrinsics)
        origin_h = torch.as_tensor(origin).detach().cpu().double().reshape(3).contiguous()
        res = float(resolution)
        X, Y, Z = tsdf_volume.shape
        tsdf_c, wvol_c = tsdf_volume.contiguous(), weights_volume.contiguous()
        world_in = None if world is None else world.detach().float().reshape(N, 3).contiguous()

        vals = torch.empty((1, N, P), dtype=torch.float32, device=dev)
        wts = torch.empty((1, N, P), dtype=torch.float32, device=dev)
        pcl = torch.empty((1, N, 3), dtype=torch.float32, device=dev)
        ray = torch.empty((N, 6), dtype=torch.float64, device=dev)
        L = _lib.lib()

        def run(o_vals, o_wts, points=None, indices=None, weights=None):
            with torch.cuda.device(dev), _lib.timed('extract' if points is None else 'extract_full', dev):
                _lib.check(L.ojdf_extract(
                    _lib.ptr(depth_f), _lib.ptr(world_in), h, w, _lib.ptr(Kinv), _lib.ptr(E), _lib.ptr(origin_h), res,
                    _lib.ptr(tsdf_c), _lib.ptr(wvol_c), X, Y, Z, P, _lib.ptr(o_vals), _lib.ptr(o_wts), _lib.ptr(pcl),
                    _lib.ptr(ray), _lib.ptr(points), _lib.ptr(indices), _lib.ptr(weights), _lib.stream_ptr(dev)))

        def materialise():
            points = torch.empty((1, N, P, 3), dtype=torch.float64, device=dev)
            indices = torch.empty((1, N, P, 8, 3), dtype=torch.int64, device=dev)
            weights = torch.empty((1, N, P, 8), dtype=torch.float64, device=dev)
            # geometry only depends on depth and pose; the gathered values go to scratch so a late
            # access (after the integrator changed the volumes) cannot disturb fusion_values/weights
            run(torch.empty_like(vals), torch.empty_like(wts), points, indices, weights)
            return dict(points=points, indices=indices, weights=weights)

        self.depth = depth_f.view(b, N)
        values = ExtractedValues(dict(fusion_values=vals, fusion_weights=wts, depth=self.depth, pcl=pcl, ray=ray),
                                 None)
        run(vals, wts)
        if eager:
            values.update(materialise())
        else:
            values._materialise = materialise
        return values
