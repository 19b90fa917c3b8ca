"""Inference engines for AdapNet++ on libojdf's tensor-core tap-GEMM kernels.

`AdapNetEngine` (bottom of this file) runs the whole network pixel-major: the 7x7 stem + BatchNorm + ReLU + max-pool is
one own kernel (ojdf_adapnet_stem), the two skip-join gates are ojdf_adapnet_skip_join, everything else except the
(optional) bilinear aux heads is a fused conv + BatchNorm + activation (+ residual) launch of ojdf_conv_tc_batched, which
picks csrc/ojdf_conv_wt.cu (>= 64 output channels on maps up to 256 pixels wide: channels as M, the image as N),
ojdf_conv_ss.cu (3x3 with a halo box that fits shared memory) or ojdf_conv_tc.cu; the two skip SSMAs run on a side
stream next to layer3 / layer4 / eASPP; stride-2
layers are the stride-1 layer followed by a consumer that reads every other pixel (`in_step = 2`); the three transposed
convolutions are 4 / 16 phase convolutions that write every 2nd / 4th output pixel (`out_step`, dead taps masked); the
SSMA gate multiplies inside the epilogue of its sigmoid convolution; `segment()` adds the fused softmax / max / arg-max
kernel.  It embeds `EncoderTailEngine`, described next.

Low-resolution tail:

At 240x320 the encoder runs layer3[1:], layer4 and the eASPP head on 15x20 feature maps with 256..2048 channels
(modules/adapnet.py:103-149,152-216).  That is 72 % of each encoder's FLOPs, and exactly the regime where the library's
fp32 convolutions collapse (300 pixels cannot fill 148 SMs: 2.5 TFLOP/s measured).  Here these layers split their K loop
over CTAs, BatchNorm / bias / ReLU / the residual add are fused into the epilogue, the conv2a|conv2b and eASPP
concatenations are channel offsets, eASPP's pooled branch is folded into a bias (computed on a side stream), and both
encoders (RGB + depth) are batched into the same launches.  `EncoderTailEngine` can also run alone (AdapNet.whole_engine =
False): the front and the decoder then stay on the module's torch forward and two tiny transposes hand the (C,15,20)
tensor over.

The eval-time-active bottleneck dropout of the reference (modules/adapnet.py:80-82) is applied between launches with
torch's own dropout, so its random stream is the library's in both paths.
Eval mode + no_grad only; rebuilt when parameters may have changed (modules/_engine_cache.py).
"""
import functools

import torch
from torch.nn import functional as F

from .. import _lib

import ctypes as C

from .fusion_engine import ConvProblem, _Conv as _ConvBase, _pad4


class StemProblem(C.Structure):
    """include/ojdf.h: ojdf_stem_problem."""
    _fields_ = [('in_dev', C.c_void_p), ('weights_dev', C.c_void_p), ('scale_dev', C.c_void_p), ('shift_dev', C.c_void_p),
                ('out_dev', C.c_void_p), ('out_stride', C.c_int), ('out_coffset', C.c_int)]

_TAIL_PIXELS = 300          # 15 x 20: four 128-pixel M-tiles per problem


def deconv_phase_weights(weight, stride, padding):
    """ConvTranspose2d(kernel 2s, stride s, padding s/2) as s*s phase convolutions (pure host logic).

    weight: (cin, cout, k, k) of the transposed convolution.  Output pixel (s*y + a, s*x + b) is a 3x3 window of
    the input around (y, x): in[y + dy, x + dx] meets kernel tap ky = a + p - s*dy, kx = b + p - s*dx when that tap
    exists (2 per axis, 4 of the 9).  Returns [(a, b, w3x3 (cout, cin, 3, 3), tap_mask)] in phase order a*s + b;
    bit t = (dy+1)*3 + (dx+1) of tap_mask is set for the live taps."""
    cin, cout, k, _ = weight.shape
    s, p = int(stride), int(padding)
    assert k == 2 * s and 2 * p == s
    phases = []
    for a in range(s):
        for b in range(s):
            wp, mask = torch.zeros(cout, cin, 3, 3, dtype=weight.dtype), 0
            for dy in (-1, 0, 1):
                ky = a + p - s * dy
                for dx in (-1, 0, 1):
                    kx = b + p - s * dx
                    if 0 <= ky < k and 0 <= kx < k:
                        wp[:, :, dy + 1, dx + 1] = weight[:, :, ky, kx].t()
                        mask |= 1 << ((dy + 1) * 3 + dx + 1)
            phases.append((a, b, wp, mask))
    return phases


def _Conv(conv, bn, act, device, n_problems=2, n_tiles=4, **kw):
    return _ConvBase(conv, bn, act, device, npad_req=0, **kw)


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

    def __init__(self, encoders, aspps, h, w, device, out_buf=None, flags=0):
        """encoders: [Encoder, ...] (1 or 2), aspps: matching eASPP modules; h, w: tail resolution.
        out_buf: optional (h*w, n*256) buffer the eASPP outputs are written into side by side."""
        self.h, self.w, self.N, self.device = int(h), int(w), int(h) * int(w), torch.device(device)
        self.flags = int(flags)
        self.side = None
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
        self.scratch = torch.zeros(self.SCRATCH_BYTES // 4, dtype=torch.float32, device=dev)   # zero: the first 16 KB are the split-K slice counters
        self.partial = torch.empty(self.PARTIAL_BLOCKS * 2048, dtype=torch.float32, device=dev)
        self._keep += [T1, T2, D]
        self.plan = []

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
        self.out = [z(co) for _ in range(n)] if out_buf is None else None
        self._keep += [cat, U]
        ms = _pad4(mid)
        # eASPP's pooled branch -> bias of the final conv: thin reductions only the LAST conv of the head needs; they run on a
        # side stream next to the branch convolutions (also inside a captured CUDA graph: a fork / join of the capture)
        self.plan.append(('fork_bias',))
        for e in range(n):
            self.plan.append(('bias', heads[e], feat[e], cmax))
        conv_step([(heads[e].b1, heads[e].b1.problem(feat[e], cmax, cat[e], 4 * co, 0)) for e in range(n)])
        conv_step([(heads[e].br[b][0], heads[e].br[b][0].problem(feat[e], cmax, U[e][b][0], ms)) for e in range(n) for b in range(3)])
        conv_step([(heads[e].br[b][1], heads[e].br[b][1].problem(U[e][b][0], ms, U[e][b][1], ms)) for e in range(n) for b in range(3)])
        conv_step([(heads[e].br[b][2], heads[e].br[b][2].problem(U[e][b][1], ms, U[e][b][0], ms)) for e in range(n) for b in range(3)])
        conv_step([(heads[e].br[b][3], heads[e].br[b][3].problem(U[e][b][0], ms, cat[e], 4 * co, (b + 1) * co))
                   for e in range(n) for b in range(3)])
        self.plan.append(('join_bias',))
        if out_buf is None:
            conv_step([(heads[e].fin, heads[e].fin.problem(cat[e], 4 * co, self.out[e], co, 0, shift=heads[e].frame_shift))
                       for e in range(n)])
        else:
            conv_step([(heads[e].fin, heads[e].fin.problem(cat[e], 4 * co, out_buf, n * co, e * co, shift=heads[e].frame_shift))
                       for e in range(n)])
        self.cmax, self.cout = cmax, co

    def run(self, st):
        """Walk the plan on the current stream (`st` = its handle); X[e][0] must hold the (N, 1024) pixel-major input of
        encoder e."""
        L = _lib.lib()
        N, H, W = self.N, self.h, self.w
        main = torch.cuda.current_stream(self.device)
        if self.side is None:
            self.side = torch.cuda.Stream(device=self.device)
        side = self.side
        for step in self.plan:
            kind = step[0]
            if kind == 'conv':
                _, arr, n, cin, cout, taps, act, slope, npad_req = step
                _lib.check(L.ojdf_conv_tc_batched(arr, n, cin, cout, H, W, taps, act, slope, 1.0, npad_req, self.flags,
                                                  self.scratch.data_ptr(), self.scratch.numel() * 4, st))
            elif kind == 'dropout':
                for t in step[1]:                            # reference quirk: active in eval mode
                    t.copy_(F.dropout(t, p=0.5, training=True))
            elif kind == 'fork_bias':
                side.wait_stream(main)
            elif kind == 'join_bias':
                main.wait_stream(side)
            else:
                _, a, src, ss = step
                _lib.check(L.ojdf_gap_bias(src.data_ptr(), ss, N, a.cin, a.wg.data_ptr(), a.g_scale.data_ptr(),
                                           a.g_shift.data_ptr(), a.cout, 1, a.wf5.data_ptr(), a.fin.scale.data_ptr(),
                                           a.fin.shift.data_ptr(), a.cout, self.partial.data_ptr(), self.PARTIAL_BLOCKS,
                                           a.frame_shift.data_ptr(), side.cuda_stream))

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
            self.run(st)
            for e in range(self.n):
                o = torch.empty(1, self.cout, H, W, dtype=torch.float32, device=dev)
                _lib.check(L.ojdf_nhwc_to_nchw(self.out[e].data_ptr(), self.cout, 0, self.cout, N, o.data_ptr(), st))
                outs.append(o)
        return outs


class AdapNetEngine:
    """Whole-network launch plan (stage 1 or 2) at input size (h, w), h and w multiples of 16."""

    def __init__(self, net, h, w, device):
        assert h % 16 == 0 and w % 16 == 0, 'AdapNet++ needs h, w = 0 mod 16 (modules/adapnet.py:288)'
        self.h, self.w, self.device = int(h), int(w), torch.device(device)
        dev = self.device
        self.net = net
        self.stage2 = net.stage != 1
        encs = [net.encoder_mod1] + ([net.encoder_mod2] if self.stage2 else [])
        aspps = [net.eASPP_mod1, net.eASPP_mod2] if self.stage2 else [net.eASPP]
        n = self.n = len(encs)
        (H4, W4), (H8, W8), (H16, W16) = (h // 4, w // 4), (h // 8, w // 8), (h // 16, w // 16)
        N4, N8, N16 = H4 * W4, H8 * W8, H16 * W16
        self.dims = (H4, W4, H8, W8, H16, W16)
        self._keep, self.plan = [], []
        self._join_w = {}
        self.side = None
        plan = self.plan

        def z(npix, c):
            """Every buffer the plan points at by raw address stays owned by the engine."""
            t = torch.zeros(npix, c, dtype=torch.float32, device=dev)
            self._keep.append(t)
            return t

        def tiles(H, W):
            return ((H + 7) // 8) * ((W + 15) // 16)

        def conv_step(pairs, H, W):
            c0 = pairs[0][0]
            arr = (ConvProblem * len(pairs))(*[p for _, p in pairs])
            self._keep.append([c for c, _ in pairs])
            plan.append(('conv', arr, len(pairs), c0.cin, c0.cout, H, W, c0.taps, c0.act, c0.slope, c0.npad_req))

        def mk(conv, bn, act, H, W, n_problems=n, **kw):
            return _Conv(conv, bn, act, dev, n_problems=n_problems, n_tiles=tiles(H, W), **kw)

        def unit(ms, src, src_stride, Hin, Win, dst, dst_stride):
            """One residual unit (torchvision Bottleneck or BottleneckSSMA) of every encoder in lock step.
            src/dst: per-encoder buffers.  A stride-2 unit computes its 3x3 at the input resolution and lets
            conv3 / the shortcut read every other pixel."""
            m0 = ms[0]
            ssma = hasattr(m0, 'conv2a')
            stride = 1 if ssma else int(m0.conv2.stride[0])
            Ho, Wo = Hin // stride, Win // stride
            Nin = Hin * Win
            mid = m0.conv1.out_channels
            c1 = [mk(m.conv1, m.bn1, 'relu', Hin, Win) for m in ms]
            T1 = [z(Nin, mid) for _ in ms]
            conv_step([(c1[e], c1[e].problem(src[e], src_stride, T1[e], mid)) for e in range(n)], Hin, Win)
            if ssma:
                half = m0.conv2a.out_channels
                c2 = [[mk(m.conv2a, m.bn2a, 'relu', Hin, Win, 2 * n), mk(m.conv2b, m.bn2b, 'relu', Hin, Win, 2 * n)] for m in ms]
                T2 = [z(Nin, 2 * half) for _ in ms]
                conv_step([(c2[e][k], c2[e][k].problem(T1[e], mid, T2[e], 2 * half, k * half)) for e in range(n) for k in range(2)],
                          Hin, Win)
                mid2 = 2 * half
            else:
                c2 = [mk(m.conv2, m.bn2, 'relu', Hin, Win) for m in ms]
                T2 = [z(Nin, mid) for _ in ms]
                conv_step([(c2[e], c2[e].problem(T1[e], mid, T2[e], mid)) for e in range(n)], Hin, Win)
                mid2 = mid
            step_kw = dict(in_step=2, in_width=Win) if stride == 2 else {}
            cout = m0.conv3.out_channels
            if m0.downsample is not None:
                assert int(m0.downsample[0].stride[0]) == stride
                dn = [mk(m.downsample[0], m.downsample[1], 'none', Ho, Wo) for m in ms]
                D = [z(Ho * Wo, cout) for _ in ms]
                conv_step([(dn[e], dn[e].problem(src[e], src_stride, D[e], cout, **step_kw)) for e in range(n)], Ho, Wo)
                res, res_stride = D, cout
            else:
                assert stride == 1
                res, res_stride = src, src_stride
            c3 = [mk(m.conv3, m.bn3, 'relu', Ho, Wo) for m in ms]
            conv_step([(c3[e], c3[e].problem(T2[e], mid2, dst[e], dst_stride, 0, residual=res[e], residual_stride=res_stride,
                                             **step_kw)) for e in range(n)], Ho, Wo)
            self._keep += [T1, T2, res]
            if ssma and m0.dropout:
                plan.append(('dropout', [dst[e][:, :cout] for e in range(n)]))
            return Ho, Wo

        # ---- encoders: stem (library) -> layer1 -> skip2 -> layer2 -> skip1 -> layer3[0] -> tail engine
        self.S0 = [z(N4, 64) for _ in range(n)]
        self.stem = []                                          # (weights (147, 64), scale, shift) of conv1 + bn1 per encoder
        for e in encs:
            r = e.res_n50_enc
            assert tuple(r.conv1.kernel_size) == (7, 7) and tuple(r.conv1.stride) == (2, 2) and tuple(r.conv1.padding) == (3, 3)
            assert r.conv1.in_channels == 3 and r.conv1.out_channels == 64 and r.conv1.bias is None
            wst = r.conv1.weight.detach().double().cpu().permute(1, 2, 3, 0).reshape(147, 64).float().contiguous().to(dev)
            sc = r.bn1.weight.detach().double().cpu() / torch.sqrt(r.bn1.running_var.detach().double().cpu() + r.bn1.eps)
            sh = r.bn1.bias.detach().double().cpu() - r.bn1.running_mean.detach().double().cpu() * sc
            self.stem.append((wst, sc.float().contiguous().to(dev), sh.float().contiguous().to(dev)))
        self.seg_scores = torch.empty(1, h, w, dtype=torch.float32, device=dev)
        self.seg_ids = torch.empty(1, h, w, dtype=torch.uint8, device=dev)
        self.seg_frame = torch.empty(1, h, w, dtype=torch.float32, device=dev)
        self.FX = z(N16, n * 256)                              # eASPP outputs of all encoders side by side
        self.flags = int(getattr(net, 'conv_flags', 0))
        self.tail = EncoderTailEngine(encs, aspps, H16, W16, dev, out_buf=self.FX, flags=self.flags)
        cur, cs = self.S0, 64
        for ui in range(len(encs[0].res_n50_enc.layer1)):
            ms = [e.res_n50_enc.layer1[ui] for e in encs]
            dst = [z(N4, 256) for _ in range(n)]
            unit(ms, cur, cs, H4, W4, dst, 256)
            cur, cs = dst, 256
        self.L1 = cur
        self.SK2 = z(N4, n * 24)                               # skip connections of all encoders side by side
        sk2 = [mk(e.enc_skip2_conv, e.enc_skip2_conv_bn, 'none', H4, W4) for e in encs]
        conv_step([(sk2[e], sk2[e].problem(cur[e], cs, self.SK2, n * 24, e * 24)) for e in range(n)], H4, W4)
        Hc, Wc = H4, W4
        for ui in range(len(encs[0].res_n50_enc.layer2)):
            ms = [e.res_n50_enc.layer2[ui] for e in encs]
            st = 1 if hasattr(ms[0], 'conv2a') else int(ms[0].conv2.stride[0])
            dst = [z((Hc // st) * (Wc // st), 512) for _ in range(n)]
            Hc, Wc = unit(ms, cur, cs, Hc, Wc, dst, 512)
            cur, cs = dst, 512
        assert (Hc, Wc) == (H8, W8)
        self.L2 = cur
        self.SK1 = z(N8, n * 24)
        sk1 = [mk(e.enc_skip1_conv, e.enc_skip1_conv_bn, 'none', H8, W8) for e in encs]
        conv_step([(sk1[e], sk1[e].problem(cur[e], cs, self.SK1, n * 24, e * 24)) for e in range(n)], H8, W8)
        self._after_skips = len(plan)                            # both skip tensors exist from here on
        ms = [e.res_n50_enc.layer3[0] for e in encs]
        unit(ms, cur, cs, H8, W8, [self.tail.X[e][0] for e in range(n)], self.tail.cmax)
        plan.append(('tail',))

        # ---- SSMA fusion of the two modalities (stage 2): cat -> 3x3 relu -> 3x3 sigmoid -> gate -> 3x3 + BN
        def ssma(m, cat, N, H, W, feat):
            red = m.link[0].out_channels
            # the gate multiplies the concatenated features (modules/adapnet.py:352): epilogue mode sigmoid(v) * residual
            l0, l1 = mk(m.link[0], None, 'relu', H, W, 1), mk(m.link[2], None, 'sigmoid_mul', H, W, 1)
            fin = mk(m.final_conv[0], m.final_conv[1], 'none', H, W, 1)
            G, GATE, out = z(N, _pad4(red)), z(N, 2 * feat), z(N, feat)
            conv_step([(l0, l0.problem(cat, 2 * feat, G, _pad4(red)))], H, W)
            conv_step([(l1, l1.problem(G, _pad4(red), GATE, 2 * feat, residual=cat, residual_stride=2 * feat))], H, W)
            conv_step([(fin, fin.problem(GATE, 2 * feat, out, feat))], H, W)
            return out

        def deconv(m, bn, act, src, src_stride, Hin, Win, dst, dst_stride):
            """ConvTranspose2d(k = 2s, stride s, padding s/2) + BatchNorm (+ act) as s*s phase convolutions:
            output pixel (s*y + a, s*x + b) = 3x3 window of the input around (y, x) with the kernel taps
            ky = a + p - s*dy, kx = b + p - s*dx that exist (4 of the 9); each phase is one problem of the
            tensor-core kernel writing every s-th pixel of the output (out_step), dead taps masked."""
            sdc, pad = int(m.stride[0]), int(m.padding[0])
            assert int(m.output_padding[0]) == 0
            Wt = m.weight.detach().cpu()                                  # (cin, cout, k, k)
            cin, cout = Wt.shape[:2]
            Wout = Win * sdc
            pairs = []
            for a, b, wp, mask in deconv_phase_weights(Wt, sdc, pad):
                fake = torch.nn.Conv2d(cin, cout, 3, padding=1, bias=True)
                fake.weight.data.copy_(wp)
                fake.bias.data.copy_(m.bias.detach().cpu() if m.bias is not None else torch.zeros(cout))
                c = mk(fake, bn, act, Hin, Win, sdc * sdc)
                pairs.append((c, c.problem(src, src_stride, dst, dst_stride, 0, out_step=sdc, out_width=Wout,
                                           tap_mask=mask, dst_ptr_offset=(a * Wout + b) * dst_stride)))
            for i in range(0, len(pairs), 8):
                conv_step(pairs[i:i + 8], Hin, Win)

        d = net.decoder
        C = int(d.n_classes)
        for conv in (d.fuse_conv1, d.fuse_conv2):               # gate of the skip joins: (24, 256) weights + bias
            assert conv.out_channels == 24 and conv.in_channels == 256
            self._join_w[id(conv)] = (conv.weight.detach().reshape(24, 256).float().contiguous().to(dev),
                                      conv.bias.detach().float().contiguous().to(dev))
        if self.stage2:
            # the two skip SSMAs (six small launches of 10..40 CTAs) only need the skip tensors: they run on a side stream
            # next to layer3 / layer4 / eASPP (whose grids leave SMs free) and join before the decoder reads them
            mark = len(plan)
            self.skip2 = ssma(net.ssma_s2, self.SK2, N4, H4, W4, 24)
            self.skip1 = ssma(net.ssma_s1, self.SK1, N8, H8, W8, 24)
            side_block = [('side_begin',)] + plan[mark:] + [('side_end',)]
            del plan[mark:]
            plan[self._after_skips:self._after_skips] = side_block
            self.X16 = ssma(net.ssma_res, self.FX, N16, H16, W16, 256)
            plan.append(('join_side',))
        else:
            self.skip2, self.skip1, self.X16 = self.SK2, self.SK1, self.FX
        # ---- decoder: transposed convolutions / aux heads stay on the library, on channels-last views
        self.J1, self.J2 = z(N8, 280), z(N4, 280)
        deconv(d.deconv1, d.deconv1_bn, 'relu', self.X16, 256, H16, W16, self.J1, 280)
        plan.append(('join1',))
        s2a, s2b = mk(d.stage2[0], d.stage2[1], 'relu', H8, W8, 1), mk(d.stage2[3], d.stage2[4], 'relu', H8, W8, 1)
        self.U1, self.U2 = z(N8, 256), z(N8, 256)
        conv_step([(s2a, s2a.problem(self.J1, 280, self.U1, 256))], H8, W8)
        conv_step([(s2b, s2b.problem(self.U1, 256, self.U2, 256))], H8, W8)
        deconv(d.stage2[6], d.stage2[7], 'none', self.U2, 256, H8, W8, self.J2, 280)
        plan.append(('join2',))
        s3a, s3b = mk(d.stage3[0], d.stage3[1], 'relu', H4, W4, 1), mk(d.stage3[3], d.stage3[4], 'relu', H4, W4, 1)
        s3c = mk(d.stage3[6], d.stage3[7], 'none', H4, W4, 1)
        Cp = _pad4(C)
        self.V1, self.V2, self.V3 = z(N4, 256), z(N4, 256), z(N4, Cp)
        self.Cp = Cp
        conv_step([(s3a, s3a.problem(self.J2, 280, self.V1, 256))], H4, W4)
        conv_step([(s3b, s3b.problem(self.V1, 256, self.V2, 256))], H4, W4)
        conv_step([(s3c, s3c.problem(self.V2, 256, self.V3, Cp))], H4, W4)
        self.LG = z(h * w, Cp)                                   # logits, pixel-major
        deconv(d.stage3[8], d.stage3[9], 'none', self.V3, Cp, H4, W4, self.LG, Cp)

    @staticmethod
    def _nchw(buf, H, W, c0=0, c1=None):
        """(1, C, H, W) channels-last view of channels [c0, c1) of a pixel-major buffer."""
        v = buf.view(1, H, W, buf.shape[1])
        return v[..., c0:c1].permute(0, 3, 1, 2)

    def forward(self, mod1, mod2=None):
        net, d, dev = self.net, self.net.decoder, self.device
        H4, W4, H8, W8, H16, W16 = self.dims
        L = _lib.lib()
        aux = {}
        with torch.cuda.device(dev), _lib.timed('adapnet_engine', dev):
            st = _lib.stream_ptr(dev)
            # stem: conv1 7x7/2 + BatchNorm + ReLU + max-pool 3x3/2 of every encoder in one launch, pixel-major out
            xs = [x.detach().float().contiguous() for x in [mod1, mod2][:self.n]]
            arr = (StemProblem * self.n)(*[StemProblem(xs[e].data_ptr(), self.stem[e][0].data_ptr(), self.stem[e][1].data_ptr(),
                                                       self.stem[e][2].data_ptr(), self.S0[e].data_ptr(), 64, 0) for e in range(self.n)])
            _lib.check(L.ojdf_adapnet_stem(arr, self.n, self.h, self.w, st))
            main = torch.cuda.current_stream(dev)
            on_side = False
            for step in self.plan:
                kind = step[0]
                if kind == 'conv':
                    _, arr, n, cin, cout, H, W, taps, act, slope, npad_req = step
                    if on_side:                                  # no split-K scratch on the side stream: it belongs to the main one
                        _lib.check(L.ojdf_conv_tc_batched(arr, n, cin, cout, H, W, taps, act, slope, 1.0, npad_req, 1 | self.flags,
                                                          None, 0, self.side.cuda_stream))
                    else:
                        _lib.check(L.ojdf_conv_tc_batched(arr, n, cin, cout, H, W, taps, act, slope, 1.0, npad_req, 1 | self.flags,
                                                          self.tail.scratch.data_ptr(), self.tail.scratch.numel() * 4, st))
                elif kind == 'side_begin':
                    if self.side is None:
                        self.side = torch.cuda.Stream(device=dev)
                    self.side.wait_stream(main)
                    on_side = True
                elif kind == 'side_end':
                    on_side = False
                elif kind == 'join_side':
                    main.wait_stream(self.side)
                elif kind == 'tail':
                    self.tail.run(st)
                elif kind == 'dropout':
                    for t in step[1]:                            # reference quirk: active in eval mode
                        t.copy_(F.dropout(t, p=0.5, training=True))
                elif kind == 'join1':
                    x = self._nchw(self.J1, H8, W8, 0, 256)      # relu(bn(deconv1(.))), written by the phase convolutions
                    aux['y1'] = d._aux(x, d.aux_conv1, d.aux_conv1_bn, 8) if net.aux_heads else None
                    self._join(x, self.skip1, d.fuse_conv1, self.J1, H8, W8)
                else:                                            # join2
                    x = self._nchw(self.J2, H4, W4, 0, 256)
                    aux['y2'] = d._aux(x, d.aux_conv2, d.aux_conv2_bn, 4) if net.aux_heads else None
                    self._join(x, self.skip2, d.fuse_conv2, self.J2, H4, W4)
            res = self._nchw(self.LG, self.h, self.w, 0, int(d.n_classes))     # (1, C, h, w) view of the pixel-major logits
        return [res, aux['y1'], aux['y2']]

    def segment(self, mod1, mod2=None):
        """forward() + the per-pixel softmax maximum / arg-max of the main head (modules/pipeline.py:57-58,184) in one
        extra launch: returns (scores (1,h,w) f32, ids (1,h,w) u8, (1 + id) / n_classes (1,h,w) f32) -- buffers owned by the
        engine, overwritten by the next call."""
        self.forward(mod1, mod2)
        C_ = int(self.net.decoder.n_classes)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ojdf_softmax_max(self.LG.data_ptr(), self.Cp, C_, self.h * self.w, C_, self.seg_scores.data_ptr(),
                                                   self.seg_ids.data_ptr(), self.seg_frame.data_ptr(), _lib.stream_ptr(self.device)))
        return self.seg_scores, self.seg_ids, self.seg_frame

    def _join(self, x, skip, conv, J, H, W):
        """Decoder._join (modules/adapnet.py:305-315): [x | gate * skip] into the 280-channel buffer J; channels [0, 256)
        already hold x.  The gate (global mean -> 1x1 conv -> ReLU) and the multiply are two launches of libojdf."""
        d = self.net.decoder
        if d.fusion:
            w, b = self._join_w[id(conv)]
            _lib.check(_lib.lib().ojdf_adapnet_skip_join(J.data_ptr(), 280, 256, H * W, w.data_ptr(), b.data_ptr(), 24, skip.data_ptr(), 24,
                                                         J.data_ptr() + 4 * 256, 280, self.tail.partial.data_ptr(), self.tail.PARTIAL_BLOCKS,
                                                         _lib.stream_ptr(self.device)))
        else:
            J.view(1, H, W, 280)[..., 256:].copy_(skip.view(1, H, W, 24))
