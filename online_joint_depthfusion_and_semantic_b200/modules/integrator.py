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
    optional semantics (N) u8 + scores (N) f32 per pixel, optional `plan` (Integrator.plan(): the geometry half of the
    step was already issued, forward() only applies the network output)."""


class IntegrationPlan:
    """Handle of an issued ojdf_integrate_plan: the workspace now holds this frame's entries grouped by voxel."""

    def __init__(self, N, P, tail, shape, workspace, event, stream):
        self.N, self.P, self.tail, self.shape, self.workspace, self.event, self.stream = N, P, tail, shape, workspace, event, stream


class Integrator(torch.nn.Module):

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.device = config.SETTINGS.device
        self.implementation = config.SETTINGS.implementation
        self._workspace = None
        self._side = None

    def plan(self, ray, filtered_depth, P, tail, volume_shape, side_stream=True):
        """Issue the geometry half of the integration (count / offsets / scatter: needs only the per-ray records and the
        masked depth) -- on a side stream by default, so it overlaps whatever the caller enqueues next on the current
        stream (the networks).  forward() with the returned plan in the FrameUpdate then runs only the apply kernels.
        The side stream first waits for the current stream: the records are ready and the previous frame's apply (which
        used the same workspace) is done."""
        _lib.require_cuda(ray, filtered_depth)
        dev = ray.device
        N = ray.shape[0]
        X, Y, Z = (int(v) for v in volume_shape)
        filt = filtered_depth.detach().float().reshape(N).contiguous()
        if N * int(tail) == 0:
            return IntegrationPlan(N, int(P), int(tail), (X, Y, Z), None, None, None)
        ws = self._get_workspace(N * int(tail) * 8, dev)
        cur = torch.cuda.current_stream(dev)
        st = cur
        if side_stream:
            if self._side is None or self._side.device != dev:
                self._side = torch.cuda.Stream(device=dev)
            st = self._side
            st.wait_stream(cur)
        with torch.cuda.device(dev), torch.cuda.stream(st), _lib.timed('integrate_plan', dev):
            _lib.check(_lib.lib().ojdf_integrate_plan(_lib.ptr(ray), _lib.ptr(filt), N, int(P), int(tail), X, Y, Z, ws.data_ptr(),
                                                      ws.numel(), st.cuda_stream))
        ev = None
        if side_stream:
            ev = torch.cuda.Event()
            ev.record(st)
            ray.record_stream(st)
            filt.record_stream(st)
        return IntegrationPlan(N, int(P), int(tail), (X, Y, Z), ws, ev, st)

    def _get_workspace(self, entries, dev):
        L = _lib.lib()
        need = int(L.ojdf_integrate_workspace_bytes(int(entries)))
        if need == 0:
            raise _lib.OjdfError('ojdf: cannot size an integration workspace for %d entries (must be 1 .. 2^31 - 2049: '
                                 'entries are indexed with 32 bits)' % entries)
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
            if N * tail == 0:                                       # empty frame: nothing to integrate
                return values_volume, weights_volume, semantics_volume, scores_volume
            plan = updates.get('plan')
            args = (N, P, tail, float(updates['clamp']), values_volume.data_ptr(), weights_volume.data_ptr(), X, Y, Z,
                    _lib.ptr(ids), _lib.ptr(sc), semantics_volume.data_ptr() if do_sem else None,
                    scores_volume.data_ptr() if do_sem else None, int(do_sem))
            if plan is not None:
                if (plan.N, plan.P, plan.tail, plan.shape) != (N, P, tail, (X, Y, Z)) or plan.workspace is None:
                    raise ValueError('integration plan was made for another frame / volume shape')
                ws = plan.workspace
                if plan.event is not None:
                    torch.cuda.current_stream(dev).wait_event(plan.event)
                with torch.cuda.device(dev), _lib.timed('integrate', dev):
                    _lib.check(L.ojdf_integrate_apply(_lib.ptr(est), *args, ws.data_ptr(), ws.numel(), stream))
            else:
                ws = self._get_workspace(N * tail * 8, dev)
                with torch.cuda.device(dev), _lib.timed('integrate', dev):
                    _lib.check(L.ojdf_integrate(_lib.ptr(ray), _lib.ptr(filt), _lib.ptr(est), *args, ws.data_ptr(), ws.numel(),
                                                stream))
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
