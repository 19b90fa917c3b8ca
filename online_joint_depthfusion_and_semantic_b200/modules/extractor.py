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

    def rays(self, depth, extrinsics, intrinsics, origin, resolution, world=None):
        """First half of forward(): world points `pcl` (1,N,3) f32 and the per-ray records `ray` (N,6) f64 (voxel-space
        centre + unit direction, modules/extractor.py:82-120,309-318).  They depend only on depth and pose, so the pipeline
        computes them before the volumes are final and plans the integration with them (Integrator.plan)."""
        b, h, w = depth.shape
        if b != 1:
            raise ValueError('the fusion path is batch-size-1 (modules/pipeline.py:199); got b=%d' % b)
        _lib.require_cuda(depth, world)
        dev = depth.device
        depth_f = depth.detach().float().contiguous()
        N = h * w
        Kinv, E = host_pose(extrinsics, intrinsics)
        origin_h = torch.as_tensor(origin).detach().cpu().double().reshape(3).contiguous()
        world_in = None if world is None else world.detach().float().reshape(N, 3).contiguous()
        pcl = torch.empty((1, N, 3), dtype=torch.float32, device=dev)
        ray = torch.empty((N, 6), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev), _lib.timed('rays', dev):
            _lib.check(_lib.lib().ojdf_rays(_lib.ptr(depth_f), _lib.ptr(world_in), h, w, _lib.ptr(Kinv), _lib.ptr(E),
                                            _lib.ptr(origin_h), float(resolution), _lib.ptr(pcl), _lib.ptr(ray), _lib.stream_ptr(dev)))
        return dict(ray=ray, pcl=pcl, depth=depth_f.view(b, N), shape=(h, w))

    def forward(self, depth, extrinsics, intrinsics, tsdf_volume, weights_volume, origin, resolution,
                world=None, eager=False, rays=None, pack=None):
        """`world` (optional, (b,N,3) f32): use these world points instead of unprojecting
        `depth` (parity tests pass the oracle's `pcl`).  `eager=True` materialises
        points/indices/weights immediately, like the reference always does.
        `rays` (optional): the result of rays() for this frame -- only the gather runs.
        `pack` (optional): (buf_a, buf_b | None, last_a, last_b | None, stride) -- FusionNet's pixel-major input buffers;
        the kernel also writes [values | weights | last] rows there (FusionNetEngine.forward(packed=True))."""
        b, h, w = depth.shape
        if b != 1:
            raise ValueError('the fusion path is batch-size-1 (modules/pipeline.py:199); got b=%d' % b)
        _lib.require_cuda(depth, tsdf_volume, weights_volume, world)
        if tsdf_volume.dtype != torch.float16 or weights_volume.dtype != torch.float16:
            raise TypeError('volumes must be float16 (modules/database.py:60,64)')
        dev = depth.device
        N, P = h * w, int(self.n_points)
        X, Y, Z = tsdf_volume.shape
        tsdf_c, wvol_c = tsdf_volume.contiguous(), weights_volume.contiguous()
        if rays is None:
            depth_f = depth.detach().float().contiguous()
            Kinv, E = host_pose(extrinsics, intrinsics)
            origin_h = torch.as_tensor(origin).detach().cpu().double().reshape(3).contiguous()
            res = float(resolution)
            world_in = None if world is None else world.detach().float().reshape(N, 3).contiguous()
            pcl = torch.empty((1, N, 3), dtype=torch.float32, device=dev)
            ray = torch.empty((N, 6), dtype=torch.float64, device=dev)
            self.depth = depth_f.view(b, N)
        else:
            assert rays['shape'] == (h, w)
            pcl, ray, self.depth = rays['pcl'], rays['ray'], rays['depth']

        vals = torch.empty((1, N, P), dtype=torch.float32, device=dev)
        wts = torch.empty((1, N, P), dtype=torch.float32, device=dev)
        L = _lib.lib()

        def run(o_vals, o_wts, points=None, indices=None, weights=None, pk=None):
            with torch.cuda.device(dev), _lib.timed('extract' if points is None else 'extract_full', dev):
                if rays is None and pk is None:
                    _lib.check(L.ojdf_extract(
                        _lib.ptr(depth_f), _lib.ptr(world_in), h, w, _lib.ptr(Kinv), _lib.ptr(E), _lib.ptr(origin_h), res,
                        _lib.ptr(tsdf_c), _lib.ptr(wvol_c), X, Y, Z, P, _lib.ptr(o_vals), _lib.ptr(o_wts), _lib.ptr(pcl),
                        _lib.ptr(ray), _lib.ptr(points), _lib.ptr(indices), _lib.ptr(weights), _lib.stream_ptr(dev)))
                    return
                if rays is None:                             # records first (same stream), then the packing gather
                    _lib.check(L.ojdf_rays(_lib.ptr(depth_f), _lib.ptr(world_in), h, w, _lib.ptr(Kinv), _lib.ptr(E),
                                           _lib.ptr(origin_h), res, _lib.ptr(pcl), _lib.ptr(ray), _lib.stream_ptr(dev)))
                pa, pb, la, lb, stride = pk if pk is not None else (None, None, None, None, 0)
                _lib.require_cuda(pa, pb, la, lb)
                _lib.check(L.ojdf_gather(_lib.ptr(ray), h, w, _lib.ptr(tsdf_c), _lib.ptr(wvol_c), X, Y, Z, P, _lib.ptr(o_vals),
                                         _lib.ptr(o_wts), _lib.ptr(points), _lib.ptr(indices), _lib.ptr(weights),
                                         _lib.ptr(pa), _lib.ptr(pb), _lib.ptr(la), _lib.ptr(lb), int(stride), _lib.stream_ptr(dev)))

        def materialise():
            points = torch.empty((1, N, P, 3), dtype=torch.float64, device=dev)
            indices = torch.empty((1, N, P, 8, 3), dtype=torch.int64, device=dev)
            weights = torch.empty((1, N, P, 8), dtype=torch.float64, device=dev)
            # geometry only depends on depth and pose; the gathered values go to scratch so a late
            # access (after the integrator changed the volumes) cannot disturb fusion_values/weights
            with torch.cuda.device(dev), _lib.timed('extract_full', dev):
                _lib.check(L.ojdf_gather(_lib.ptr(ray), h, w, _lib.ptr(tsdf_c), _lib.ptr(wvol_c), X, Y, Z, P,
                                         _lib.ptr(torch.empty_like(vals)), _lib.ptr(torch.empty_like(wts)), _lib.ptr(points),
                                         _lib.ptr(indices), _lib.ptr(weights), None, None, None, None, 0, _lib.stream_ptr(dev)))
            return dict(points=points, indices=indices, weights=weights)

        values = ExtractedValues(dict(fusion_values=vals, fusion_weights=wts, depth=self.depth, pcl=pcl, ray=ray),
                                 None)
        run(vals, wts, pk=pack)
        if eager:
            values.update(materialise())
        else:
            values._materialise = materialise
        return values
