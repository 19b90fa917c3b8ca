"""Inference engine for the low-resolution tail of AdapNet++ on libojdf's tap-GEMM kernels.

At 240x320 the encoder runs layer3[1:], layer4 and the eASPP head on 15x20 feature maps with 256..2048
channels (modules/adapnet.py:103-149,152-216).  That is 72 % of each encoder's FLOPs, and exactly the
regime where the library's fp32 convolutions collapse (300 pixels cannot fill 148 SMs: 2.5 TFLOP/s
measured).  Here these layers run pixel-major on `conv_tile_kernel` with the K loop split across
blocks, BatchNorm / bias / ReLU / the residual add fused into the epilogue, the conv2a|conv2b and
eASPP concatenations as channel offsets, eASPP's pooled branch folded into a bias, and both
encoders (RGB + depth) batched into the same launches.

Everything before (conv1 .. layer3[0], the two skip convs) and after (SSMA, decoder) stays on the
module's torch forward; two tiny transposes hand the (C,15,20) tensor over.  The eval-time-active
bottleneck dropout of the reference (modules/adapnet.py:80-82) is applied between launches with
torch's own dropout, so its random stream is the library's in both paths.
Eval mode + no_grad only; rebuilt when parameters may have changed (same rules as FusionNetEngine).
"""
import functools

import torch
from torch.nn import functional as F

from .. import _lib

from .fusion_engine import ConvProblem, _Conv as _ConvBase, _pad4, conv_mode

_TAIL_PIXELS = 300          # 15 x 20: four 128-pixel M-tiles per problem


def _npad_for(cout, n_problems, n_tiles=4):
    """Output-channel group width of the tensor-core kernel on the small tail maps: the widest of
    128 / 64 / 32 that still yields ~one CTA per SM (tiles x groups x problems >= 120)."""
    for npad in (128, 64, 32):
        if n_tiles * n_problems * ((cout + npad - 1) // npad) >= 120:
            return npad
    return 32


def _Conv(conv, bn, act, device, n_problems=2, **kw):
    tc = conv_mode() == 'tc'
    return _ConvBase(conv, bn, act, device, tc=tc, npad_req=_npad_for(conv.out_channels, n_problems) if tc else 0, **kw)


class _Unit:
    """Bottleneck (torchvision) or multi-scale unit (BottleneckSSMA) as fused conv specs."""

    def __init__(self, m, device):
        self.ssma = hasattr(m, 'conv2a')
        self.c1 = _Conv(m.conv1, m.bn1, 'relu', device)
        if self.ssma:
            self.c2 = [_Conv(m.conv2a, m.bn2a, 'relu', device, n_problems=4), _Conv(m.conv2b, m.bn2b, 'relu', device, n_problems=4)]
            self.dropout = bool(m.dropout)
        else:
            self.c2 = [_Conv(m.conv2, m.bn2, 'relu', device)]
            self.dropout = False
        for c in [m.conv1, m.conv3] + ([m.conv2a, m.conv2b] if self.ssma else [m.conv2]):
            assert tuple(c.stride) == (1, 1), 'the 15x20 tail has no strided convolution'
        self.c3 = _Conv(m.conv3, m.bn3, 'relu', device)                  # ReLU after the residual add
        self.down = _Conv(m.downsample[0], m.downsample[1], 'none', device) if m.downsample is not None else None
        if m.downsample is not None:
            assert tuple(m.downsample[0].stride) == (1, 1)
        self.cin, self.cmid, self.cout = self.c1.cin, self.c1.cout, self.c3.cout


class _ASPP:
    def __init__(self, m, device):
        self.b1 = _Conv(m.branch1_conv, m.branch1_bn, 'relu', device)
        self.br = [[_Conv(b[0], b[1], 'relu', device, n_problems=6), _Conv(b[3], b[4], 'relu', device, n_problems=6),
                    _Conv(b[6], b[7], 'relu', device, n_problems=6), _Conv(b[9], b[10], 'relu', device, n_problems=6)]
                   for b in m.branch234]
        co = self.b1.cout
        self.cout, self.cin, self.mid = co, self.b1.cin, self.br[0][0].cout
        self.fin = _Conv(m.eASPP_fin_conv, m.eASPP_fin_bn, 'relu', device, cin_slice=(0, 4 * co))    # branches 1-4
        # branch 5: relu(conv(mean)) (no BN, modules/adapnet.py:203-204) -> bias of the final conv
        self.wg = m.branch5_conv.weight.detach().reshape(co, self.cin).float().contiguous().to(device)
        self.g_scale = torch.ones(co, dtype=torch.float32, device=device)
        self.g_shift = m.branch5_conv.bias.detach().float().contiguous().to(device)
        self.wf5 = m.eASPP_fin_conv.weight.detach()[:, 4 * co:5 * co].reshape(co, co).float().contiguous().to(device)
        self.frame_shift = torch.empty(co, dtype=torch.float32, device=device)


class EncoderTailEngine:
    SCRATCH_BYTES = 96 << 20
    PARTIAL_BLOCKS = 296

    def __init__(self, encoders, aspps, h, w, device):
        """encoders: [Encoder, ...] (1 or 2), aspps: matching eASPP modules; h, w: tail resolution."""
        self.h, self.w, self.N, self.device = int(h), int(w), int(h) * int(w), torch.device(device)
        dev, N, n = self.device, self.N, len(encoders)
        self.n = n
        units = [[_Unit(m, dev) for m in list(e.res_n50_enc.layer3)[1:] + list(e.res_n50_enc.layer4)] for e in encoders]
        heads = [_ASPP(a, dev) for a in aspps]
        self._keep = [units, heads]
        z = lambda c: torch.zeros(N, c, dtype=torch.float32, device=dev)     # noqa: E731
        cmax = max(u.cout for u in units[0])
        self.cin0 = units[0][0].cin
        self.X = [[z(cmax), z(cmax)] for _ in range(n)]                       # ping-pong unit input / output
        T1, T2, D = [z(512) for _ in range(n)], [z(512) for _ in range(n)], [z(cmax) for _ in range(n)]
        self.scratch = torch.empty(self.SCRATCH_BYTES // 4, dtype=torch.float32, device=dev)
        self.partial = torch.empty(self.PARTIAL_BLOCKS * 2048, dtype=torch.float32, device=dev)
        self._keep += [T1, T2, D]
        self.plan = []
        self.tc = conv_mode() == 'tc'

        def conv_step(pairs):
            c0 = pairs[0][0]
            arr = (ConvProblem * len(pairs))(*[p for _, p in pairs])
            self.plan.append(('conv', arr, len(pairs), c0.cin, c0.cout, c0.taps, c0.act, c0.slope, c0.npad_req))

        cur = 0
        for ui in range(len(units[0])):
            us = [units[e][ui] for e in range(n)]
            u0 = us[0]
            src = [self.X[e][cur] for e in range(n)]
            dst = [self.X[e][cur ^ 1] for e in range(n)]
            conv_step([(us[e].c1, us[e].c1.problem(src[e], cmax, T1[e], 512)) for e in range(n)])
            half = u0.c2[0].cout
            conv_step([(us[e].c2[k], us[e].c2[k].problem(T1[e], 512, T2[e], 512, k * half))
                       for e in range(n) for k in range(len(u0.c2))])
            if u0.down is not None:
                conv_step([(us[e].down, us[e].down.problem(src[e], cmax, D[e], cmax)) for e in range(n)])
                res = D
            else:
                res = src
            conv_step([(us[e].c3, us[e].c3.problem(T2[e], 512, dst[e], cmax, 0, residual=res[e], residual_stride=cmax))
                       for e in range(n)])
            if u0.dropout:
                self.plan.append(('dropout', [dst[e][:, :u0.cout] for e in range(n)]))
            cur ^= 1
        feat = [self.X[e][cur] for e in range(n)]
        A = heads[0]
        co, mid = A.cout, A.mid
        cat = [z(4 * co) for _ in range(n)]
        U = [[[z(_pad4(mid)) for _ in range(2)] for _ in range(3)] for _ in range(n)]
        self.out = [z(co) for _ in range(n)]
        self._keep += [cat, U]
        ms = _pad4(mid)
        for e in range(n):
            self.plan.append(('bias', heads[e], feat[e], cmax))
        conv_step([(heads[e].b1, heads[e].b1.problem(feat[e], cmax, cat[e], 4 * co, 0)) for e in range(n)])
        conv_step([(heads[e].br[b][0], heads[e].br[b][0].problem(feat[e], cmax, U[e][b][0], ms)) for e in range(n) for b in range(3)])
        conv_step([(heads[e].br[b][1], heads[e].br[b][1].problem(U[e][b][0], ms, U[e][b][1], ms)) for e in range(n) for b in range(3)])
        conv_step([(heads[e].br[b][2], heads[e].br[b][2].problem(U[e][b][1], ms, U[e][b][0], ms)) for e in range(n) for b in range(3)])
        conv_step([(heads[e].br[b][3], heads[e].br[b][3].problem(U[e][b][0], ms, cat[e], 4 * co, (b + 1) * co))
                   for e in range(n) for b in range(3)])
        conv_step([(heads[e].fin, heads[e].fin.problem(cat[e], 4 * co, self.out[e], co, 0, shift=heads[e].frame_shift))
                   for e in range(n)])
        self.cmax, self.cout = cmax, co

    def forward(self, xs):
        """xs: list of (1, C, h, w) NCHW tensors (output of layer3[0] of each encoder).
        Returns a list of (1, 256, h, w) NCHW tensors (eASPP outputs)."""
        L = _lib.lib()
        dev, N, H, W = self.device, self.N, self.h, self.w
        outs = []
        with torch.cuda.device(dev), _lib.timed('adapnet_tail', dev):
            st = _lib.stream_ptr(dev)
            for e, x in enumerate(xs):
                x = x.detach().float().contiguous()
                _lib.check(L.ojdf_nchw_to_nhwc(x.data_ptr(), self.cin0, N, self.X[e][0].data_ptr(), self.cmax, 0, st))
            for step in self.plan:
                kind = step[0]
                if kind == 'conv':
                    _, arr, n, cin, cout, taps, act, slope, npad_req = step
                    if self.tc:
                        _lib.check(L.ojdf_conv_tc_batched(arr, n, cin, cout, H, W, taps, act, slope, 1.0, npad_req, 0, st))
                    else:
                        _lib.check(L.ojdf_conv_nhwc_batched(arr, n, cin, cout, H, W, taps, act, slope, 1.0,
                                                            self.scratch.data_ptr(), self.scratch.numel() * 4, st))
                elif kind == 'dropout':
                    for t in step[1]:                            # reference quirk: active in eval mode
                        t.copy_(F.dropout(t, p=0.5, training=True))
                else:
                    _, a, src, ss = step
                    _lib.check(L.ojdf_gap_bias(src.data_ptr(), ss, N, a.cin, a.wg.data_ptr(), a.g_scale.data_ptr(),
                                               a.g_shift.data_ptr(), a.cout, 1, a.wf5.data_ptr(), a.fin.scale.data_ptr(),
                                               a.fin.shift.data_ptr(), a.cout, self.partial.data_ptr(), self.PARTIAL_BLOCKS,
                                               a.frame_shift.data_ptr(), st))
            for e in range(self.n):
                o = torch.empty(1, self.cout, H, W, dtype=torch.float32, device=dev)
                _lib.check(L.ojdf_nhwc_to_nchw(self.out[e].data_ptr(), self.cout, 0, self.cout, N, o.data_ptr(), st))
                outs.append(o)
        return outs
