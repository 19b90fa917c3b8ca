"""FusionNet_v2 / FusionNet_v3 -- host-side mirror of the reference's modules/model.py.

The parameter tree (and therefore every state_dict key, e.g. `block0.0.block.0.weight`,
`vortex0.branches.2.3.weight`, `vortex3.final.0.weight`, `pred.4.pred.6.bias`) is identical to
the reference's so checkpoints load in both directions (test_fusion.py:63-65):

  Block   (modules/model.py:4-21)    conv3x3-BN-LeakyReLU-Drop2d(0.2) x2, params at seq idx 0,1,4,5
  Pred    (modules/model.py:24-52)   1x1 conv chain; the last one ends conv-LReLU-conv-tanh (idx 0,1,4,6)
  VortexPooling (modules/model.py:100-161)  GAP branch (idx 1,3) + 4 dilated branches (rates 1,3,9,27;
                                     idx 0,1,3,4,6,7,9,10) fed by cascaded 3x3 average pools + 1x1 `final`
  FusionNet_v3 (modules/model.py:219-283)   two dense-block heads (TSDF / semantics) -> vortex -> vortex3 -> pred

The layer containers are built from compact specs rather than written out, and forward() is
expressed over a small functional core (`_run`) so that the sm_100a conv kernels can take the
place of the library convolutions layer group by layer group without touching the parameters.
FusionNet_v1 is not provided: it cannot be constructed in the reference either
(modules/model.py:58 raises NameError).
"""
import torch
from torch import nn

from ._engine_cache import EngineOwner

_VORTEX_RATES = (1, 3, 9, 27)


def _conv(cin, cout, k=1, dilation=1):
    return nn.Conv2d(cin, cout, kernel_size=k, padding=dilation * (k // 2), dilation=dilation)


def _sequential(spec):
    """spec: list of ('conv', cin, cout, k, dil) | ('bn', c) | 'lrelu' | 'relu' | 'drop' | 'tanh'."""
    layers = []
    for s in spec:
        if s == 'lrelu':
            layers.append(nn.LeakyReLU())
        elif s == 'relu':
            layers.append(nn.ReLU())
        elif s == 'drop':
            layers.append(nn.Dropout2d(p=0.2))
        elif s == 'tanh':
            layers.append(nn.Tanh())
        elif s[0] == 'conv':
            layers.append(_conv(*s[1:]))
        elif s[0] == 'bn':
            layers.append(nn.BatchNorm2d(s[1]))
        else:
            raise ValueError(s)
    return nn.Sequential(*layers)


class Block(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        c = out_channels
        self.block = _sequential([('conv', in_channels, c, 3, 1), ('bn', c), 'lrelu', 'drop',
                                  ('conv', c, c, 3, 1), ('bn', c), 'lrelu', 'drop'])

    def forward(self, x):
        return self.block(x)


class Pred(nn.Module):
    def __init__(self, in_channels, out_channels, n_points=None):
        super().__init__()
        c = out_channels
        head = [('conv', in_channels, c, 1, 1), ('bn', c), 'lrelu', 'drop']
        # the reference always builds the BN variant first and then replaces it for the output stage
        # (modules/model.py:30-50); doing the same keeps seeded random initialisation identical
        self.pred = _sequential(head + [('conv', c, c, 1, 1), ('bn', c), 'lrelu', 'drop'])
        if n_points is not None:
            self.pred = _sequential(head + [('conv', c, c, 1, 1), 'lrelu', ('conv', c, n_points, 1, 1), 'tanh'])

    def forward(self, x):
        return self.pred(x)


class _Broadcast1x1(nn.Module):
    """nn.Upsample(size, mode='bilinear', align_corners=True) of a 1x1 map (modules/model.py:110): every output pixel is the
    input value exactly (the interpolation weights are 1 and 0), i.e. a broadcast.  Written as expand() its backward is a
    plain sum; the library's bilinear backward instead adds all H*W gradients of a channel into ONE address with atomics
    (50 ms per call at 240x320, 3/4 of a training frame, tools/train_profile.py).  No parameters: same state_dict."""

    def __init__(self, size):
        super().__init__()
        self.size = tuple(size) if isinstance(size, (tuple, list)) else (int(size), int(size))

    def forward(self, x):
        assert x.shape[-2:] == (1, 1)
        return x.expand(-1, -1, self.size[0], self.size[1]).contiguous()


class VortexPooling(nn.Module):
    def __init__(self, in_chs, mid_chs, out_chs, feat_res):
        super().__init__()
        self.gave_pool = nn.Sequential(
            nn.AdaptiveAvgPool2d((1, 1)),
            _conv(in_chs, out_chs, 1),
            _Broadcast1x1(feat_res),
            nn.BatchNorm2d(num_features=out_chs))
        self.pool1 = nn.AvgPool2d(kernel_size=3, stride=1, padding=1)
        self.pool2 = nn.AvgPool2d(kernel_size=3, stride=1, padding=1)
        self.pool3 = nn.AvgPool2d(kernel_size=3, stride=1, padding=1)
        self.branches = nn.ModuleList([
            _sequential([('conv', in_chs, mid_chs, 1, 1), ('bn', mid_chs), 'relu',
                         ('conv', mid_chs, mid_chs, 3, r), ('bn', mid_chs), 'relu',
                         ('conv', mid_chs, mid_chs, 3, r), ('bn', mid_chs), 'relu',
                         ('conv', mid_chs, out_chs, 1, 1), ('bn', out_chs), 'relu'])
            for r in _VORTEX_RATES])
        self.final = nn.Sequential(_conv(5 * out_chs, out_chs, 1), nn.BatchNorm2d(num_features=out_chs),
                                   nn.Dropout2d(p=0.2, inplace=True))

    def forward(self, x):
        feats = [self.gave_pool(x), self.branches[0](x)]
        xp = x
        for pool, branch in zip((self.pool1, self.pool2, self.pool3), list(self.branches)[1:]):
            xp = pool(xp)                                   # cascaded: 3x3, 5x5, 7x7 effective windows
            feats.append(branch(xp))
        return self.final(torch.cat(feats, dim=1))


class _EngineMixin(EngineOwner):
    """Eval-mode, no-grad, CUDA forwards run on libojdf's fused kernels (fusion_engine.py); anything
    that needs autograd (training, row a2) runs the module's torch forward.  The launch plan is
    dropped whenever parameters can have changed: train(), .to()/.cuda(), load_state_dict() on this module
    or on a parent, in-place writes (see _engine_cache.py)."""
    _engine = None
    use_engine = True

    def _drop_engines(self):
        self._engine = None

    def train(self, mode=True):
        self._invalidate()
        return super().train(mode)

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._invalidate()
        return super().load_state_dict(*a, **k)

    def engine_ready(self, ref_tensor):
        return (self.use_engine and not self.training and not torch.is_grad_enabled() and ref_tensor.is_cuda
                and not (hasattr(self, 'block') and self.config.use_semantics))

    def engine_for(self, h, w, device):
        """The launch plan for (h, w) frames on `device` (built on first use, rebuilt when parameters changed)."""
        from .fusion_engine import FusionNetEngine
        self._hook_load_state_dict()
        self.engines_current()
        e = self._engine
        if e is None or (e.h, e.w) != (int(h), int(w)) or e.device != torch.device(device):
            e = self._engine = FusionNetEngine(self, h, w, device)
        return e

    def forward_pixel_major(self, vals, wts, frame, sem_frame=None, packed=False):
        """vals/wts (1,N,P) pixel-major, frame (1,h,w) depth, sem_frame (1,h,w) normalised labels
        -> est (1,N,P), already multiplied by output_scale.  packed=True: the input rows were already written into the
        engine's buffers (Extractor.forward(pack=engine.pack_target(...)))."""
        h, w = frame.shape[-2:]
        return self.engine_for(h, w, vals.device).forward(vals, wts, frame, sem_frame, packed=packed)

    def _forward_engine_nchw(self, x):
        """Same call surface as the reference forward (dict of NCHW tensors in, NCHW tensor out)."""
        b, P, h, w = x['tsdf_values'].shape
        pm = lambda t: t.permute(0, 2, 3, 1).reshape(b, h * w, -1)           # noqa: E731
        sem = x['semantic_frame'].reshape(b, h, w) if self.config.use_semantics else None
        est = self.forward_pixel_major(pm(x['tsdf_values']), pm(x['tsdf_weights']), x['tsdf_frame'].reshape(b, h, w), sem)
        return est.view(b, h, w, -1).permute(0, 3, 1, 2)


def _dense_blocks(x, blocks):
    for blk in blocks:
        x = torch.cat([x, blk(x)], dim=1)
    return x


def _pred_stack(n_channels, gf, n_points):
    return nn.Sequential(*[Pred((gf + 1 - i) * n_channels, (gf - i) * n_channels,
                                n_points if i == gf - 1 else None) for i in range(gf)])


class FusionNet_v2(_EngineMixin, nn.Module):
    """modules/model.py:164-216: one head on [values, weights, frame(, semantics)]."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.scale = config.output_scale
        self.n_points = config.n_points
        self.resy, self.resx = config.resy, config.resx
        self.n_channels = config.n_points * 2 + 1 + int(config.use_semantics)
        self.gf = config.growth_factor - 1
        pool_in = self.n_channels * (self.gf + 1)
        self.block = nn.ModuleList([Block((i + 1) * self.n_channels, self.n_channels) for i in range(self.gf)])
        self.vortex = VortexPooling(pool_in, self.n_channels, pool_in, (self.resy, self.resx))
        self.vortex_final = VortexPooling(pool_in, self.n_channels, pool_in, (self.resy, self.resx))
        self.pred = _pred_stack(self.n_channels, self.gf, self.n_points)

    def forward(self, x):
        if self.engine_ready(x['tsdf_values']):
            return self._forward_engine_nchw(x)
        parts = [x['tsdf_values'], x['tsdf_weights'], x['tsdf_frame']]
        if self.config.use_semantics:
            parts.append(x['semantic_frame'])
        y = _dense_blocks(torch.cat(parts, dim=1), self.block)
        y = self.vortex_final(self.vortex(y))
        return self.pred(y) * self.scale


class FusionNet_v3(_EngineMixin, nn.Module):
    """modules/model.py:219-283: TSDF head (+ semantic head) -> vortex3 -> pred."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.scale = config.output_scale
        self.n_points = config.n_points
        self.resy, self.resx = config.resy, config.resx
        self.n_channels = config.n_points * 2 + 1
        self.gf = config.growth_factor - 1
        pool_in = self.n_channels * (self.gf + 1)
        res = (self.resy, self.resx)
        heads = 1
        self.block0 = nn.ModuleList([Block((i + 1) * self.n_channels, self.n_channels) for i in range(self.gf)])
        self.vortex0 = VortexPooling(pool_in, self.n_channels, pool_in, res)
        if self.config.use_semantics:
            heads += 1
            self.block2 = nn.ModuleList([Block((i + 1) * self.n_channels, self.n_channels) for i in range(self.gf)])
            self.vortex2 = VortexPooling(pool_in, self.n_channels, pool_in, res)
        self.vortex3 = VortexPooling(heads * pool_in, self.n_channels, pool_in, res)
        self.pred = _pred_stack(self.n_channels, self.gf, self.n_points)

    def forward(self, x):
        if self.engine_ready(x['tsdf_values']):
            return self._forward_engine_nchw(x)
        y = self.vortex0(_dense_blocks(torch.cat([x['tsdf_values'], x['tsdf_weights'], x['tsdf_frame']], dim=1),
                                       self.block0))
        if self.config.use_semantics:
            y1 = self.vortex2(_dense_blocks(
                torch.cat([x['tsdf_values'], x['tsdf_weights'], x['semantic_frame']], dim=1), self.block2))
            y = torch.cat([y, y1], dim=1)
        return self.pred(self.vortex3(y)) * self.scale
