"""Pipeline -- drop-in for the reference's modules/pipeline.py:12-363.

Same construction (`Pipeline(config)`, sub-modules `_fusion_network`, `_semantic_2d_network`,
`_extractor`, `_integrator`), same driver calls (`fuse(batch, database, device) -> None`,
`fuse_training(batch, database, device) -> {'tsdf_est','tsdf_fused','tsdf_target'}`), same
database protocol (reads `database[scene]`, re-binds `scenes_est/fusion_weights/ids_est/scores`,
modules/pipeline.py:199-244) -- so test_fusion.py / train_fusion.py drive it unchanged.

Per frame (modules/pipeline.py:173-248):
  [AdapNet -> softmax -> max]  ->  Extractor  ->  FusionNet  ->  volume update  ->  Integrator
What differs from the reference is only *how* the memory-bound steps run: the extractor and the
integrator are one libojdf call each, and the update handed to the integrator is the compact
`FrameUpdate` (per-ray record + network output + masked depth) instead of ~100 MB of gathered
indices / weights / values (modules/pipeline.py:150-169).
"""
import os

import torch
from torch import nn

from .. import _lib
from ..cuda_graph import GraphedCall
from .adapnet import AdapNet
from .extractor import Extractor
from .integrator import FrameUpdate, Integrator
from .model import FusionNet_v2, FusionNet_v3

_FUSION_NETS = {'v2': FusionNet_v2, 'v3': FusionNet_v3}


class Pipeline(nn.Module):

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.n_points = config.FUSION_MODEL.n_points
        if config.DATA.semantics:
            self.n_classes = config.SEMANTIC_2D_MODEL.n_classes
        config.FUSION_MODEL.resx = config.DATA.resx
        config.FUSION_MODEL.resy = config.DATA.resy
        name = config.FUSION_MODEL.name
        if name not in _FUSION_NETS:
            raise ValueError("FUSION_MODEL.name must be 'v2' or 'v3' (the reference's 'v1' cannot be "
                             "constructed, modules/model.py:58); got %r" % (name,))
        self._fusion_network = _FUSION_NETS[name](config.FUSION_MODEL)
        if config.DATA.semantics and config.DATA.semantic_strategy == 'predict':
            self._semantic_2d_network = AdapNet(config.SEMANTIC_2D_MODEL)
        else:
            self._semantic_2d_network = None
        self._extractor = Extractor(config)
        self._integrator = Integrator(config)
        self.use_cuda_graphs = True          # replay the fixed-shape networks as CUDA graphs in inference
        # issue the geometry half of the integration on a side stream, early (OJDF_PLAN_AHEAD=0: plan + apply after FusionNet)
        self.plan_ahead = os.environ.get('OJDF_PLAN_AHEAD', '1') != '0'
        self._seg_graph = None
        self._seg_graph_token = (None, None)
        self._seg_graph_fn = None
        self._sem_frame = None

    def set_precision(self, mode):
        """'parity' (3xTF32, default) or 'fast' (1xTF32) for the tensor-core convolutions of both networks."""
        self._seg_graph = None
        self._fusion_network.set_precision(mode)
        if self._semantic_2d_network is not None:
            self._semantic_2d_network.set_precision(mode)
        return self

    def train(self, mode=True):
        self._seg_graph = None               # captured graphs hold parameter addresses / modes
        return super().train(mode)

    def _apply(self, fn, *a, **k):
        self._seg_graph = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._seg_graph = None
        return super().load_state_dict(*a, **k)

    # ---- a3: 2-D segmentation (modules/pipeline.py:42-60) -------------------------------------
    def _segmentation_eager(self, image, aux):
        """image (b,3,h,w) raw batch image, aux (b,h,w)|(b,1,h,w) second modality or None."""
        image = (image / 255.0).float()                                # quirk: /255 after mean/std normalisation
        net = self._semantic_2d_network
        net.aux_heads = False                                          # only the main head is read below (:54-58)
        if self.config.SEMANTIC_2D_MODEL.stage == 1:
            logits = net(image if aux is None else aux.repeat(1, 3, 1, 1).float())[0]
        else:
            logits = net(image, aux.repeat(1, 3, 1, 1).float())[0]
        return torch.softmax(logits, dim=1).permute(0, 2, 3, 1)

    def _seg_inputs(self, image, aux):
        image = (image / 255.0).float()                                # quirk: /255 after mean/std normalisation
        if self.config.SEMANTIC_2D_MODEL.stage == 1:
            return (image if aux is None else aux.repeat(1, 3, 1, 1).float()), None
        return image, aux.repeat(1, 3, 1, 1).float()

    def _segment_fast(self, image, aux):
        """(scores, ids u8, label frame): AdapNet++ launch plan + fused softmax / max / arg-max kernel."""
        net = self._semantic_2d_network
        net.aux_heads = False
        return net.segment(*self._seg_inputs(image, aux))

    def _graphed(self, fn, image, aux):
        """Replay `fn(image, aux)` as a CUDA graph.  The captured graph replays kernels that point into the AdapNet++ launch
        plan: it is only valid while that very plan object is alive and its parameters are unchanged (eval() / .to() /
        load_state_dict() / in-place writes on the network drop the plan, modules/_engine_cache.py)."""
        net = self._semantic_2d_network
        net._hook_load_state_dict()
        net.engines_current()
        if self._seg_graph is not None and (self._seg_graph_fn != fn or
                                            any(a is not b for a, b in zip(net.engine_token(), self._seg_graph_token))):
            self._seg_graph = None
        fresh = self._seg_graph is None
        if fresh:
            self._seg_graph = GraphedCall(lambda *a: fn(a[0], a[1] if len(a) > 1 else None))
            self._seg_graph_fn = fn
        out = self._seg_graph(image) if aux is None else self._seg_graph(image, aux)
        if fresh:
            self._seg_graph_token = net.engine_token()
        return out

    def _segmentation(self, data):
        key = self.config.DATA.input
        image = data['image'].to(self.device)
        aux = None if key == 'image' else data[key].to(self.device)
        net = self._semantic_2d_network
        graphable = (self.use_cuda_graphs and image.is_cuda and not torch.is_grad_enabled() and not net.training)
        if not graphable:
            return self._segmentation_eager(image, aux)
        return self._graphed(self._segmentation_eager, image, aux)

    def segment(self, batch, device=None):
        """The 2-D segmentation of a frame on its own (modules/pipeline.py:181-184): private copies of (scores, ids u8, label
        frame) that `fuse(..., semantics=...)` accepts in place of running AdapNet++ itself.  It only needs the batch, not
        the volumes, so a streaming caller (stream.FrameStream) runs it for frame i+1 on a second stream while frame i is
        still being extracted / fused / integrated.  None when the pipeline does not predict semantics."""
        if not (self.config.DATA.semantics and self.config.DATA.semantic_strategy == 'predict'):
            return None
        if device is not None:
            self.device = device
        self._shape = batch['image'].shape
        scores, ids = self._semantic_frame(batch, as_uint8=True)
        frame = self._sem_frame
        return {'scores': scores.clone(), 'ids': ids.clone(), 'sem_frame': None if frame is None else frame.clone()}

    def _semantic_frame(self, batch, as_uint8):
        """(scores f32, ids) per pixel, or (None, None): modules/pipeline.py:181-193,277-292."""
        if not self.config.DATA.semantics:
            return None, None
        pre, self._pre_semantics = getattr(self, '_pre_semantics', None), None
        if pre is not None:                                      # computed ahead by segment()
            self._sem_frame = pre['sem_frame']
            return pre['scores'], (pre['ids'] if as_uint8 else pre['ids'].long())
        strategy = self.config.DATA.semantic_strategy
        self._sem_frame = None
        if strategy == 'predict':
            net = self._semantic_2d_network
            image = batch['image'].to(self.device)
            key = self.config.DATA.input
            aux = None if key == 'image' else batch[key].to(self.device)
            with torch.no_grad(), _lib.timed('adapnet', self.device):      # AdapNet++ is frozen on the fusion path
                if net.whole_engine_ready(image):
                    fn = self._segment_fast
                    scores, ids, self._sem_frame = (self._graphed(fn, image, aux) if self.use_cuda_graphs else fn(image, aux))
                    return scores, (ids if as_uint8 else ids.long())
                scores, ids = self._segmentation(batch).max(dim=-1)
        elif strategy == 'gt':
            ids = batch['semantic_gt'].to(self.device).long()
            scores = torch.ones_like(ids).float()
        else:
            raise ValueError('Error! Valid value for DATA.semantic_strategy are "gt" or "predict".')
        return scores, (ids.type(torch.uint8) if as_uint8 else ids)

    # ---- a9/a11: network input / output plumbing (modules/pipeline.py:62-102) -------------------
    def _prepare_fusion_input(self, frame, values, semantics):
        b, _, h, w = self._shape
        P = self.n_points
        inputs = {
            'tsdf_values': values['fusion_values'].view(b, h, w, P),
            'tsdf_weights': values['fusion_weights'].view(b, h, w, P),
            'tsdf_frame': frame.unsqueeze(-1),
        }
        if self.config.FUSION_MODEL.use_semantics:
            assert semantics is not None
            inputs['semantic_frame'] = (1 + semantics.unsqueeze(-1).float()) / self.n_classes      # (0, 1]
        return {k: v.permute(0, 3, 1, 2).contiguous() for k, v in inputs.items()}

    def _fusion(self, inputs, values, packed=False):
        b, _, h, w = self._shape
        net = self._fusion_network
        if getattr(net, 'engine_ready', None) is not None and net.engine_ready(values['fusion_values']):
            # pixel-major in, pixel-major out: the extractor's layout is the kernels' layout
            sem = inputs['semantic_frame'].reshape(b, h, w) if 'semantic_frame' in inputs else None
            return net.forward_pixel_major(values['fusion_values'], values['fusion_weights'],
                                           inputs['tsdf_frame'].reshape(b, h, w), sem, packed=packed)[..., :self.n_points]
        est = net(inputs).permute(0, 2, 3, 1)[..., :self.n_points]
        return est.reshape(b, h * w, self.n_points)

    def _pack_target(self, frame, sem_ids):
        """FusionNet's own input buffers, if its launch plan will run this frame: the extractor's gather then writes the
        [values | weights | depth or label] rows itself (modules/pipeline.py:74-102 without the packing pass)."""
        net = self._fusion_network
        if getattr(net, 'engine_ready', None) is None or not net.engine_ready(frame):
            return None
        b, _, h, w = self._shape
        sem = None
        if self.config.FUSION_MODEL.use_semantics:
            assert sem_ids is not None
            sem = self._sem_frame if self._sem_frame is not None else (1 + sem_ids.reshape(b, h, w).float()) / self.n_classes
        return net.engine_for(h, w, frame.device).pack_target(frame.reshape(b, h, w), sem)

    # ---- a12: loss tensors (modules/pipeline.py:104-135) ------------------------------------------
    def _prepare_fusion_output(self, values, tsdf_est, filtered_frame=None, values_gt=None):
        b, _, h, w = self._shape
        lim = self.config.DATA.init_value
        old = values['fusion_values']
        wts = values['fusion_weights'].view(b, h * w, self.n_points).clamp(min=0)
        fused = (wts * old + torch.clamp(tsdf_est, -lim, lim)) / (wts + 1)
        if values_gt is None:
            return fused
        assert filtered_frame is not None
        keep = (filtered_frame.view(b, h * w, 1) != 0.)[0, :, 0].nonzero()[:, 0]
        target = values_gt['fusion_values'].view(b, h * w, self.n_points)
        return {'tsdf_est': tsdf_est, 'tsdf_fused': fused[:, keep], 'tsdf_target': target[:, keep]}

    # ---- a13: volume update (modules/pipeline.py:137-171), compact form ----------------------------
    def _prepare_volume_update(self, values, tsdf_est, tsdf_frame, semantics, scores):
        b, _, h, w = self._shape
        upd = FrameUpdate(ray=values['ray'], filtered_depth=tsdf_frame.reshape(h * w),
                          est=tsdf_est.detach().reshape(h * w, -1), tail=self.config.FUSION_MODEL.n_tail_points,
                          clamp=self.config.DATA.init_value)
        if self.config.DATA.semantics:
            assert semantics is not None and scores is not None
            upd['semantics'] = semantics.to(self.device).reshape(h * w)
            upd['scores'] = scores.to(self.device).reshape(h * w)
        return upd

    def _frames(self, batch):
        frame = batch[self.config.DATA.input].squeeze_(1).to(self.device)
        return frame, torch.where(batch['mask'].to(self.device), frame, torch.zeros_like(frame))

    # ---- a1: inference step (modules/pipeline.py:173-248) -------------------------------------------
    def fuse(self, batch, database, device, semantics=None):
        """`semantics`: optional result of segment(batch) (skips the AdapNet++ pass of this call)."""
        self.device = device
        self._shape = batch['image'].shape
        self._pre_semantics = semantics
        frame, filtered_frame = self._frames(batch)
        scene_id = batch['frame_id'][0].split('/')[0]
        volume = database[scene_id]
        # per-ray records first: they only need depth and pose, and with them the geometry half of the integration
        # (grouping the frame's (ray, sample, corner) entries by voxel) is issued on a side stream, where it overlaps the
        # two networks; only the light apply kernels remain after FusionNet
        rays = plan = None
        if frame.is_cuda and self.plan_ahead:
            rays = self._extractor.rays(frame, batch['extrinsics'], batch['intrinsics'], volume['origin'], volume['resolution'])
            plan = self._integrator.plan(rays['ray'], filtered_frame, self.n_points, self.config.FUSION_MODEL.n_tail_points,
                                         volume['current'].shape)
        scores, sem_ids = self._semantic_frame(batch, as_uint8=True)     # labels stay u8 end to end (no .long() / .type() round trip)
        pack = self._pack_target(frame, sem_ids) if frame.is_cuda else None
        values = self._extractor.forward(frame, batch['extrinsics'], batch['intrinsics'], volume['current'],
                                         volume['weights'], volume['origin'], volume['resolution'], rays=rays, pack=pack)
        if pack is not None:
            # the gather already wrote FusionNet's input rows: only views of the two per-pixel frames are handed over (the
            # NCHW copies of modules/pipeline.py:74-102 would be four launches nobody reads)
            inputs = {'tsdf_frame': frame.unsqueeze(1)}
            if pack[3] is not None:
                inputs['semantic_frame'] = pack[3].view(frame.shape).unsqueeze(1)
        else:
            inputs = self._prepare_fusion_input(frame, values, sem_ids)
        tsdf_est = self._fusion(inputs, values, packed=pack is not None)
        sem = self.config.DATA.semantics
        if sem:
            sem_ids = sem_ids.type(torch.uint8)
        updates = self._prepare_volume_update(values, tsdf_est, filtered_frame, sem_ids if sem else None, scores)
        if plan is not None:
            updates['plan'] = plan
        tsdf, weights, ids, sc = self._integrator.forward(updates, volume['current'], volume['weights'],
                                                          volume['scores'] if sem else None,
                                                          volume['ids_est'] if sem else None)
        database.state[scene_id] = True
        database.scenes_est[scene_id].volume = tsdf
        database.fusion_weights[scene_id] = weights
        if sem:
            database.ids_est[scene_id].volume = ids
            database.scores[scene_id].volume = sc

    # ---- a2: training step (modules/pipeline.py:251-363) ---------------------------------------------
    def fuse_training(self, batch, database, device):
        self.device = device
        self._shape = batch['image'].shape
        scores, sem_ids = self._semantic_frame(batch, as_uint8=True)
        frame, filtered_frame = self._frames(batch)
        scene_id = batch['frame_id'][0].split('/')[0]
        volume = database[scene_id]
        pose = (batch['extrinsics'], batch['intrinsics'])
        values = self._extractor.forward(frame, *pose, volume['current'], volume['weights'], volume['origin'],
                                         volume['resolution'])
        values_gt = self._extractor.forward(frame, *pose, volume['gt'], volume['weights'], volume['origin'],
                                            volume['resolution'])
        tsdf_est = self._fusion(self._prepare_fusion_input(frame, values, sem_ids), values)
        output = self._prepare_fusion_output(values, tsdf_est, filtered_frame, values_gt)
        sem = self.config.DATA.semantics
        updates = self._prepare_volume_update(values, tsdf_est, filtered_frame, sem_ids if sem else None, scores)
        # the semantic volumes are not updated while training (test=False, modules/pipeline.py:349-357)
        tsdf, weights, _, _ = self._integrator.forward(updates, volume['current'], volume['weights'],
                                                       volume['scores'] if sem else None,
                                                       volume['ids_est'] if sem else None, test=False)
        database.state[scene_id] = True
        database.scenes_est[scene_id].volume = tsdf.detach()
        database.fusion_weights[scene_id] = weights.detach()
        return output
