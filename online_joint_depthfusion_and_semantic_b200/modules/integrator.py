"""Integrator -- drop-in for the reference's modules/integrator.py:5-126.

forward(updates, values_volume, weights_volume, scores_volume, semantics_volume, test=True)
returns (values_volume, weights_volume, semantics_volume, scores_volume) -- the reference's
order (modules/integrator.py:126) -- after updating the SAME tensors in place.

`updates` is either
  * the reference's dict {values (1,Nv,T) f32, indices (1,Nv,T,8,3) i64, weights (1,Nv,T,8)
    f64[, semantics (1,Nv,T,1) u8, scores (1,Nv,T,1) f32]} (modules/pipeline.py:137-171), or
  * a `FrameUpdate`, the compact whole-frame form this package's Pipeline builds: the
    extractor's per-ray record + the network output + the masked depth, so that the 100+ MB of
    gathered indices/weights are never written.
Both go through libojdf's deterministic scatter/finalize kernels (bit-identical to the
single-threaded reference).  CUDA only; no fallback.
"""
import torch

from .. import _lib


class FrameUpdate(dict):
    """Whole-frame update: ray (N,6) f64, filtered_depth (N) f32, est (N,P) f32, tail, clamp,
    optional semantics (N) u8 + scores (N) f32 per pixel."""


class Integrator(torch.nn.Module):

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.device = config.SETTINGS.device
        self.implementation = config.SETTINGS.implementation
        self._workspace = None

    def _get_workspace(self, entries, dev):
        L = _lib.lib()
        need = int(L.ojdf_integrate_workspace_bytes(int(entries)))
        if need == 0:
            raise _lib.OjdfError('ojdf: %d entries do not fit the 32-bit entry index' % entries)
        ws = self._workspace
        if ws is None or ws.numel() < need or ws.device != dev:
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                _lib.check(L.ojdf_integrate_workspace_init(ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)))
            self._workspace = ws
        return ws

    def forward(self, updates, values_volume, weights_volume, scores_volume, semantics_volume, test=True):
        _lib.require_cuda(values_volume, weights_volume)
        if values_volume.dtype != torch.float16 or weights_volume.dtype != torch.float16:
            raise TypeError('TSDF / weight volumes must be float16 (modules/database.py:60,64)')
        if not (values_volume.is_contiguous() and weights_volume.is_contiguous()):
            raise ValueError('volumes are updated in place and must be contiguous')
        dev = values_volume.device
        X, Y, Z = values_volume.shape
        do_sem = bool(self.config.DATA.semantics) and bool(test)
        if do_sem:
            _lib.require_cuda(scores_volume, semantics_volume)
            if semantics_volume.dtype != torch.uint8 or scores_volume.dtype != torch.float16:
                raise TypeError('label volume must be uint8 and score volume float16 (modules/database.py:70,74)')
            if not (semantics_volume.is_contiguous() and scores_volume.is_contiguous()):
                raise ValueError('volumes are updated in place and must be contiguous')
        L = _lib.lib()
        stream = _lib.stream_ptr(dev)

        if isinstance(updates, FrameUpdate):
            ray = updates['ray']
            N = ray.shape[0]
            est = updates['est'].detach().float().reshape(N, -1).contiguous()
            P, tail = est.shape[1], int(updates['tail'])
            filt = updates['filtered_depth'].detach().float().reshape(N).contiguous()
            ids = sc = None
            if do_sem:
                ids = updates['semantics'].detach().reshape(N).to(torch.uint8).contiguous()
                sc = updates['scores'].detach().reshape(N).float().contiguous()
            ws = self._get_workspace(N * tail * 8, dev)
            with torch.cuda.device(dev), _lib.timed('integrate', dev):
                _lib.check(L.ojdf_integrate(
                    _lib.ptr(ray), _lib.ptr(filt), _lib.ptr(est), N, P, tail, float(updates['clamp']),
                    values_volume.data_ptr(), weights_volume.data_ptr(), X, Y, Z, _lib.ptr(ids), _lib.ptr(sc),
                    semantics_volume.data_ptr() if do_sem else None, scores_volume.data_ptr() if do_sem else None,
                    int(do_sem), ws.data_ptr(), ws.numel(), stream))
        else:
            values = updates['values'].to(dev).detach().float().contiguous()
            M1 = values.numel()
            indices = updates['indices'].to(dev).detach().long().reshape(M1, 8, 3).contiguous()
            weights = updates['weights'].to(dev).detach().double().reshape(M1, 8).contiguous()
            ids = sc = None
            if do_sem:
                ids = updates['semantics'].to(dev).detach().reshape(M1).to(torch.uint8).contiguous()
                sc = updates['scores'].to(dev).detach().reshape(M1).float().contiguous()
            if M1 > 0:
                ws = self._get_workspace(M1 * 8, dev)
                with torch.cuda.device(dev), _lib.timed('integrate_updates', dev):
                    _lib.check(L.ojdf_integrate_updates(
                        _lib.ptr(values), _lib.ptr(indices), _lib.ptr(weights), M1,
                        values_volume.data_ptr(), weights_volume.data_ptr(), X, Y, Z, _lib.ptr(ids), _lib.ptr(sc),
                        semantics_volume.data_ptr() if do_sem else None, scores_volume.data_ptr() if do_sem else None,
                        int(do_sem), ws.data_ptr(), ws.numel(), stream))

        return values_volume, weights_volume, semantics_volume, scores_volume
