"""Inference engine for FusionNet_v2 / FusionNet_v3 on libojdf's fp32 tap-GEMM kernels.

The nn.Module in model.py owns the parameters (so checkpoints load exactly like the reference's,
test_fusion.py:63-65); this class turns them into a launch plan over pixel-major (NHWC) buffers:

  * per conv: weights re-laid out as [out-group of 20][tap][cin padded to 4][20], conv bias +
    inference BatchNorm folded into a per-channel (scale, shift) epilogue, activation fused;
  * dense-block concatenation (modules/model.py:258-263) = channel offsets into one buffer;
  * VortexPooling (modules/model.py:100-161): 3 cascaded 3x3 average pools, 4 dilated branches writing
    into one 4*C buffer, the global-pool branch folded into the bias of the `final` 1x1 conv;
  * Pred chain (modules/model.py:24-52) ends in tanh * output_scale and writes (N, n_points) f32 --
    exactly the layout the integrator consumes, so no NCHW<->NHWC permutes exist anywhere.

Only used in eval mode under torch.no_grad(); training (autograd through FusionNet, row a2) keeps
the module's own torch forward.  The plan is rebuilt when parameters or buffers change.
"""
import torch
from torch import nn

from .. import _lib

_GROUP = 20
_ACT = {'none': 0, 'relu': 1, 'lrelu': 2, 'tanh': 3}


def _pad4(c):
    return (c + 3) // 4 * 4


class _Conv:
    """One fused conv (+BN) (+activation) launch."""

    def __init__(self, conv, bn, act, device, cin_slice=None, slope=0.01):
        w = conv.weight.detach().double()                       # (cout, cin, kh, kw)
        if cin_slice is not None:
            w = w[:, cin_slice[0]:cin_slice[1]]
        cout, cin, kh, kw = w.shape
        assert kh == kw and kh in (1, 3)
        self.taps, self.cin, self.cout = kh * kw, cin, cout
        self.dil = int(conv.dilation[0])
        assert kh == 1 or int(conv.padding[0]) == self.dil
        groups, cin_p = (cout + _GROUP - 1) // _GROUP, _pad4(cin)
        prep = torch.zeros(groups, self.taps, cin_p, _GROUP, dtype=torch.float64)
        wt = w.permute(2, 3, 1, 0).reshape(self.taps, cin, cout)             # [tap][ci][co], tap = ky*3+kx
        for g in range(groups):
            n = min(_GROUP, cout - g * _GROUP)
            prep[g, :, :cin, :n] = wt[:, :, g * _GROUP:g * _GROUP + n]
        bias = conv.bias.detach().double() if conv.bias is not None else torch.zeros(cout, dtype=torch.float64)
        if bn is not None:
            s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
            t = (bias - bn.running_mean.detach().double()) * s + bn.bias.detach().double()
        else:
            s, t = torch.ones(cout, dtype=torch.float64), bias
        self.weights = prep.float().contiguous().to(device)
        self.scale = s.float().contiguous().to(device)
        self.shift = t.float().contiguous().to(device)
        self.act, self.slope = _ACT[act], float(slope)

    def run(self, L, stream, H, W, src, src_stride, dst, dst_stride, dst_off=0, out_mul=1.0, shift=None):
        _lib.check(L.ojdf_conv_nhwc(src.data_ptr(), src_stride, self.cin, H, W, self.taps, self.dil,
                                    self.weights.data_ptr(), self.scale.data_ptr(),
                                    (self.shift if shift is None else shift).data_ptr(), self.cout,
                                    self.act, self.slope, float(out_mul), dst.data_ptr(), dst_stride, dst_off, stream))


class _Vortex:
    def __init__(self, m, device):
        gp_conv, gp_bn = m.gave_pool[1], m.gave_pool[3]
        self.cin, self.cout = gp_conv.in_channels, gp_conv.out_channels
        self.mid = m.branches[0][0].out_channels
        self.branches = []
        for br in m.branches:
            self.branches.append([_Conv(br[0], br[1], 'relu', device), _Conv(br[3], br[4], 'relu', device),
                                  _Conv(br[6], br[7], 'relu', device), _Conv(br[9], br[10], 'relu', device)])
        fin_conv, fin_bn = m.final[0], m.final[1]
        C = self.cout
        self.final = _Conv(fin_conv, fin_bn, 'none', device, cin_slice=(C, 5 * C))     # the 4 branch outputs
        # global branch: v1 = BN(conv(mean)); its share of the final conv becomes a bias
        self.wg = gp_conv.weight.detach().reshape(C, self.cin).float().contiguous().to(device)
        gb = gp_conv.bias.detach().double()
        gs = gp_bn.weight.detach().double() / torch.sqrt(gp_bn.running_var.detach().double() + gp_bn.eps)
        self.g_scale = gs.float().to(device)
        self.g_shift = ((gb - gp_bn.running_mean.detach().double()) * gs + gp_bn.bias.detach().double()).float().to(device)
        self.wf1 = fin_conv.weight.detach()[:, :C].reshape(C, C).float().contiguous().to(device)
        self.frame_shift = torch.empty(C, dtype=torch.float32, device=device)


class FusionNetEngine:
    PARTIAL_BLOCKS = 296

    def __init__(self, net, h, w, device):
        self.h, self.w, self.N, self.device = int(h), int(w), int(h) * int(w), torch.device(device)
        self.P = int(net.n_points)
        self.scale = float(net.scale)
        self.v3 = hasattr(net, 'block0')
        self.use_sem = bool(net.config.use_semantics)
        self.nch = int(net.n_channels)
        self.gf = int(net.gf)
        dev, N = self.device, self.N

        def blocks(ml):
            return [(_Conv(b.block[0], b.block[1], 'lrelu', dev), _Conv(b.block[4], b.block[5], 'lrelu', dev)) for b in ml]

        C = self.nch * (self.gf + 1)                            # 114
        self.C, self.Cs = C, _pad4(C)
        z = lambda c: torch.zeros(N, c, dtype=torch.float32, device=dev)     # noqa: E731
        if self.v3:
            self.heads = [(blocks(net.block0), _Vortex(net.vortex0, dev))]
            if self.use_sem:
                self.heads.append((blocks(net.block2), _Vortex(net.vortex2, dev)))
            self.tail_vortex = [_Vortex(net.vortex3, dev)]
            self.in_bufs = [z(self.Cs) for _ in self.heads]
            self.cat = z(_pad4(len(self.heads) * C))
            self.cat_stride = _pad4(len(self.heads) * C)
        else:
            if self.use_sem:
                raise NotImplementedError('FusionNet_v2 with a semantic input channel: use the torch forward')
            self.heads = [(blocks(net.block), _Vortex(net.vortex, dev))]
            self.tail_vortex = [_Vortex(net.vortex_final, dev)]
            self.in_bufs = [z(self.Cs)]
            self.cat, self.cat_stride = z(self.Cs), self.Cs
        self.pred = []
        for pm in net.pred:
            seq = pm.pred
            if isinstance(seq[5], nn.BatchNorm2d):
                self.pred += [_Conv(seq[0], seq[1], 'lrelu', dev), _Conv(seq[4], seq[5], 'lrelu', dev)]
            else:                                              # last Pred: conv-BN-LReLU, conv-LReLU, conv-tanh
                self.pred += [_Conv(seq[0], seq[1], 'lrelu', dev), _Conv(seq[4], None, 'lrelu', dev),
                              _Conv(seq[6], None, 'tanh', dev)]
        cmax = max(v.cin for _, v in self.heads + [(None, self.tail_vortex[0])])
        self.mid_s = _pad4(self.nch)
        self.t19 = [z(self.mid_s) for _ in range(3)]
        self.pool = [z(_pad4(cmax)) for _ in range(3)]
        self.pool_stride = _pad4(cmax)
        self.branch = z(4 * C)
        self.vout = z(self.Cs)
        self.pp = [z(_pad4(C)) for _ in range(2)]
        self.partial = torch.empty(self.PARTIAL_BLOCKS * 256, dtype=torch.float32, device=dev)

    def _vortex(self, L, st, v, src, src_stride, dst, dst_stride, dst_off):
        H, W, C = self.h, self.w, v.cout
        _lib.check(L.ojdf_vortex_bias(src.data_ptr(), src_stride, self.N, v.cin, v.wg.data_ptr(), v.g_scale.data_ptr(),
                                      v.g_shift.data_ptr(), C, v.wf1.data_ptr(), v.final.scale.data_ptr(),
                                      v.final.shift.data_ptr(), C, self.partial.data_ptr(), self.PARTIAL_BLOCKS,
                                      v.frame_shift.data_ptr(), st))
        cur, cur_stride = src, src_stride
        for i, br in enumerate(v.branches):
            if i > 0:                                           # cascaded 3x3 average pools
                _lib.check(L.ojdf_avgpool3_nhwc(cur.data_ptr(), cur_stride, H, W, _pad4(v.cin), self.pool[i - 1].data_ptr(),
                                                self.pool_stride, st))
                cur, cur_stride = self.pool[i - 1], self.pool_stride
            br[0].run(L, st, H, W, cur, cur_stride, self.t19[0], self.mid_s)
            br[1].run(L, st, H, W, self.t19[0], self.mid_s, self.t19[1], self.mid_s)
            br[2].run(L, st, H, W, self.t19[1], self.mid_s, self.t19[2], self.mid_s)
            br[3].run(L, st, H, W, self.t19[2], self.mid_s, self.branch, 4 * C, i * C)
        v.final.run(L, st, H, W, self.branch, 4 * C, dst, dst_stride, dst_off, shift=v.frame_shift)

    def forward(self, vals, wts, frame, sem_frame=None):
        """vals / wts: (1,N,P) f32 pixel-major (Extractor output), frame: (1,h,w) f32 depth,
        sem_frame: (1,h,w) f32 normalised labels (1+id)/n_classes (modules/pipeline.py:96).
        Returns est (1,N,P) f32."""
        _lib.require_cuda(vals, wts, frame, sem_frame)
        dev, N, H, W, P = self.device, self.N, self.h, self.w, self.P
        L = _lib.lib()
        vals = vals.detach().float().contiguous()
        wts = wts.detach().float().contiguous()
        frame = frame.detach().float().contiguous()
        two = self.v3 and self.use_sem
        if two:
            assert sem_frame is not None
            sem_frame = sem_frame.detach().float().contiguous()
        est = torch.empty(1, N, P, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), _lib.timed('fusionnet', dev):
            st = _lib.stream_ptr(dev)
            _lib.check(L.ojdf_pack_fusion_input(vals.data_ptr(), wts.data_ptr(), frame.data_ptr(),
                                                sem_frame.data_ptr() if two else None, N, P, self.in_bufs[0].data_ptr(),
                                                self.in_bufs[1].data_ptr() if two else None, self.Cs, st))
            for hi, (blocks, vortex) in enumerate(self.heads):
                buf = self.in_bufs[hi]
                for bi, (c1, c2) in enumerate(blocks):
                    c1.run(L, st, H, W, buf, self.Cs, self.t19[0], self.mid_s)
                    c2.run(L, st, H, W, self.t19[0], self.mid_s, buf, self.Cs, (bi + 1) * self.nch)
                self._vortex(L, st, vortex, buf, self.Cs, self.cat, self.cat_stride, hi * self.C)
            self._vortex(L, st, self.tail_vortex[0], self.cat, self.cat_stride, self.vout, self.Cs, 0)
            cur, cur_stride = self.vout, self.Cs
            for i, c in enumerate(self.pred):
                last = i == len(self.pred) - 1
                dst, dst_stride = (est, P) if last else (self.pp[i % 2], self.pp[i % 2].shape[1])
                c.run(L, st, H, W, cur, cur_stride, dst, dst_stride, 0, out_mul=self.scale if last else 1.0)
                cur, cur_stride = dst, dst_stride
        return est
