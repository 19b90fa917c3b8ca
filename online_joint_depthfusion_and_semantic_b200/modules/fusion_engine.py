"""Inference engine for FusionNet_v2 / FusionNet_v3 on libojdf's tensor-core convolution kernels.

The nn.Module in model.py owns the parameters (so checkpoints load exactly like the reference's,
test_fusion.py:63-65); this class turns them into a launch plan over pixel-major (NHWC) buffers:

  * per conv: weights packed once as swizzled shared-memory images (ojdf_conv_tc_pack_weights), conv bias +
    inference BatchNorm folded on the host in f64 into a per-channel (scale, shift) epilogue, activation fused;
  * dense-block concatenation (modules/model.py:258-263) = channel offsets into one buffer;
  * VortexPooling (modules/model.py:100-161): the first 1x1 convolution of all four branches is ONE launch (the bare
    products W_b . x side by side); the 3 cascaded 3x3 average pools commute with that convolution, so the 19-channel
    products are pooled (3 batched launches, bias / BatchNorm / ReLU after the last pool; an identity problem of the
    first launch gives branch 0 its epilogue) instead of the 114-channel input; two launches of dilated 3x3 convolutions
    (dilations 1 / 3 / 9 / 27 of all heads batched); then ONE chain launch (ojdf_conv_chain) runs the four 19 -> 114
    branch-out convolutions and the `final` convolution: the concatenation becomes a sum over four 114-column slices of
    `final`, the branch outputs live only in tensor memory; the global-pool branch is folded into the bias of `final`
    (computed on a side stream);
  * Pred (modules/model.py:24-52): all eleven 1x1 layers as ONE chain launch, ending in tanh * output_scale and writing
    (N, n_points) f32 -- exactly the layout the integrator consumes, so no NCHW<->NHWC permutes exist anywhere.
    OJDF_CHAIN=0 / net.use_chain = False: one launch per layer.

Channel groups are padded to a multiple of 4 channels inside the buffers (19 -> 20, 114 -> 116): every
concatenation offset is then 16-byte aligned, which is what the kernels' TMA-store epilogue
needs; the consumers' weights carry zero rows at the pad positions (`cin_map`).

Only used in eval mode under torch.no_grad(); training (autograd through FusionNet, row a2) keeps
the module's own torch forward.  The plan is rebuilt when parameters or buffers change.
"""
import ctypes as C
import os

import numpy as np
import torch
from torch import nn

from .. import _lib

_ACT = {'none': 0, 'relu': 1, 'lrelu': 2, 'tanh': 3, 'sigmoid': 4, 'sigmoid_mul': 5}


def _pad4(c):
    return (c + 3) // 4 * 4


def group_map(n_groups, width):
    """Positions of n_groups x width real channels when every group is padded to a multiple of 4."""
    pitch = _pad4(width)
    return [g * pitch + c for g in range(n_groups) for c in range(width)], n_groups * pitch


def pack_tc_weights(w, npad_req=0):
    """(cout, cin, kh, kw) f64/f32 host tensor -> packed float32 image for ojdf_conv_tc_batched."""
    cout, cin, kh, kw = w.shape
    taps = kh * kw
    L = _lib.lib()
    src = np.ascontiguousarray(w.float().numpy().reshape(cout, cin, taps))
    out = np.zeros(L.ojdf_conv_tc_weight_floats(cin, cout, taps, npad_req), np.float32)
    _lib.check(L.ojdf_conv_tc_pack_weights(src.ctypes.data, cin, cout, taps, npad_req, out.ctypes.data))
    return torch.from_numpy(out)


class ConvProblem(C.Structure):
    """include/ojdf.h: ojdf_conv_problem."""
    _fields_ = [('in_dev', C.c_void_p), ('weights_dev', C.c_void_p), ('scale_dev', C.c_void_p), ('shift_dev', C.c_void_p),
                ('out_dev', C.c_void_p), ('residual_dev', C.c_void_p), ('in_stride', C.c_int), ('out_stride', C.c_int),
                ('out_coffset', C.c_int), ('dilation', C.c_int), ('residual_stride', C.c_int), ('in_step', C.c_int),
                ('in_width', C.c_int), ('out_step', C.c_int), ('out_width', C.c_int), ('tap_mask', C.c_int)]


class ChainInput(C.Structure):
    """include/ojdf.h: ojdf_chain_input."""
    _fields_ = [('in_dev', C.c_void_p * 2), ('in_stride', C.c_int), ('cin', C.c_int)]


class ChainStep(C.Structure):
    """include/ojdf.h: ojdf_chain_step."""
    _fields_ = [('weights_dev', C.c_void_p * 2), ('scale_dev', C.c_void_p * 2), ('shift_dev', C.c_void_p * 2), ('src', C.c_int),
                ('cin', C.c_int), ('cout', C.c_int), ('acc', C.c_int), ('fresh', C.c_int), ('epi', C.c_int), ('act', C.c_int),
                ('slope', C.c_float)]


def _ptrs(tensors):
    """Two-slot pointer array of the chain structs (one slot per problem)."""
    return (C.c_void_p * 2)(*([t.data_ptr() if t is not None else None for t in tensors] + [None] * (2 - len(tensors))))


class PoolProblem(C.Structure):
    """include/ojdf.h: ojdf_pool_problem."""
    _fields_ = [('in_dev', C.c_void_p), ('out_dev', C.c_void_p), ('scale_dev', C.c_void_p), ('shift_dev', C.c_void_p),
                ('in_stride', C.c_int), ('out_stride', C.c_int), ('identity', C.c_int)]


class _Conv:
    """One fused conv (+BN) (+activation): weights packed for the tensor-core kernels, BatchNorm folded on the host in f64."""

    def __init__(self, conv, bn, act, device, cin_slice=None, slope=0.01, cin_map=None, npad_req=0, raw=False):
        w = conv.weight.detach().double().cpu()                 # (cout, cin, kh, kw); all folding on the host in f64
        if cin_slice is not None:
            w = w[:, cin_slice[0]:cin_slice[1]]
        if cin_map is not None:                                 # real input channel i sits at position cin_map[0][i] of
            pos, width = cin_map                                # a `width`-channel padded buffer; pads get zero weights
            assert len(pos) == w.shape[1]
            wp = torch.zeros(w.shape[0], width, w.shape[2], w.shape[3], dtype=torch.float64)
            wp[:, torch.as_tensor(pos, dtype=torch.long)] = w
            w = wp
        cout, cin, kh, kw = w.shape
        assert kh == kw and kh in (1, 3)
        self.taps, self.cin, self.cout = kh * kw, cin, cout
        self.dil = int(conv.dilation[0])
        assert kh == 1 or int(conv.padding[0]) == self.dil
        bias = conv.bias.detach().double().cpu() if conv.bias is not None else torch.zeros(cout, dtype=torch.float64)
        if bn is not None:
            s = bn.weight.detach().double().cpu() / torch.sqrt(bn.running_var.detach().double().cpu() + bn.eps)
            t = (bias - bn.running_mean.detach().double().cpu()) * s + bn.bias.detach().double().cpu()
        else:
            s, t = torch.ones(cout, dtype=torch.float64), bias
        if raw:                                                # the bare product W.x: bias / BatchNorm / activation applied later
            s, t, act = torch.ones(cout, dtype=torch.float64), torch.zeros(cout, dtype=torch.float64), 'none'
        self.npad_req = int(npad_req)
        self.weights_tc = pack_tc_weights(w, self.npad_req).to(device)
        self.scale = s.float().contiguous().to(device)
        self.shift = t.float().contiguous().to(device)
        self.act, self.slope = _ACT[act], float(slope)

    def problem(self, src, src_stride, dst, dst_stride, dst_off=0, shift=None, residual=None, residual_stride=0,
                in_step=0, in_width=0, out_step=0, out_width=0, tap_mask=0, dst_ptr_offset=0):
        return ConvProblem(src.data_ptr(), self.weights_tc.data_ptr(), self.scale.data_ptr(),
                           (self.shift if shift is None else shift).data_ptr(), dst.data_ptr() + 4 * dst_ptr_offset,
                           None if residual is None else residual.data_ptr(), src_stride, dst_stride, dst_off, self.dil,
                           residual_stride, in_step, in_width, out_step, out_width, tap_mask)


class _Vortex:
    def __init__(self, m, device, in_map):
        """in_map = (positions, padded width) of the module's input channels in the source buffer."""
        gp_conv, gp_bn = m.gave_pool[1], m.gave_pool[3]
        self.cout = gp_conv.out_channels
        self.cin = in_map[1]                                    # padded input width
        self.branches = [[_Conv(br[0], br[1], 'relu', device, cin_map=in_map), _Conv(br[3], br[4], 'relu', device),
                          _Conv(br[6], br[7], 'relu', device), _Conv(br[9], br[10], 'relu', device)] for br in m.branches]
        # branches 1..3 see pool^b(x); their first 1x1 conv commutes with the pools, so it runs on x itself without
        # bias / BatchNorm / ReLU (`raw`), and those are applied after the last pool (scale / shift below, padded to 4)
        # all four first-layer products in ONE convolution: 4 groups of (19 -> 20) output rows, bias / BatchNorm / ReLU later
        mid = m.branches[0][0].out_channels
        Gm = _pad4(mid)
        merged = nn.Conv2d(m.branches[0][0].in_channels, 4 * Gm, 1, bias=False)
        with torch.no_grad():
            merged.weight.zero_()
            for b, br in enumerate(m.branches):
                merged.weight[b * Gm:b * Gm + mid] = br[0].weight.detach().cpu()
        self.raw_all = _Conv(merged, None, 'none', device, cin_map=in_map, raw=True)
        self.post = []
        for b in range(0, 4):
            f = self.branches[b][0]
            sc, sh = torch.zeros(_pad4(f.cout), device=device), torch.zeros(_pad4(f.cout), device=device)
            sc[:f.cout], sh[:f.cout] = f.scale, f.shift
            self.post.append((sc, sh))
        fin_conv, fin_bn = m.final[0], m.final[1]
        C_ = self.cout
        # the 4 branch outputs, each in its own 4-aligned group of the branch buffer
        self.final = _Conv(fin_conv, fin_bn, 'none', device, cin_slice=(C_, 5 * C_), cin_map=group_map(4, C_))
        # the same convolution as four 114-column slices, one per branch (chain kernel: the concatenation becomes a sum)
        self.final_b = [_Conv(fin_conv, fin_bn, 'none', device, cin_slice=(C_ * (1 + b), C_ * (2 + b))) for b in range(4)]
        # global branch: v1 = BN(conv(mean)); its share of the final conv becomes a bias
        wg = torch.zeros(C_, self.cin, dtype=torch.float32)
        wg[:, torch.as_tensor(in_map[0], dtype=torch.long)] = gp_conv.weight.detach().reshape(C_, -1).float().cpu()
        self.wg = wg.contiguous().to(device)
        gb = gp_conv.bias.detach().double()
        gs = gp_bn.weight.detach().double() / torch.sqrt(gp_bn.running_var.detach().double() + gp_bn.eps)
        self.g_scale = gs.float().to(device)
        self.g_shift = ((gb - gp_bn.running_mean.detach().double()) * gs + gp_bn.bias.detach().double()).float().to(device)
        self.wf1 = fin_conv.weight.detach()[:, :C_].reshape(C_, C_).float().contiguous().to(device)
        self.frame_shift = torch.empty(C_, dtype=torch.float32, device=device)


class FusionNetEngine:
    """Launch plan built once per (network, frame size); forward() only walks it."""
    PARTIAL_BLOCKS = 296

    def __init__(self, net, h, w, device):
        self.h, self.w, self.N, self.device = int(h), int(w), int(h) * int(w), torch.device(device)
        self.P = int(net.n_points)
        self.scale = float(net.scale)
        self.v3 = hasattr(net, 'block0')
        self.use_sem = bool(net.config.use_semantics)
        nch, gf = int(net.n_channels), int(net.gf)
        dev, N = self.device, self.N
        if not self.v3 and self.use_sem:
            raise NotImplementedError('FusionNet_v2 with a semantic input channel: use the torch forward')

        G = _pad4(nch)                                          # pitch of one 19-channel group: 20
        Cc = nch * (gf + 1)                                     # 114 real channels after the dense blocks
        Cd = (gf + 1) * G                                       # ... living in 6 groups of 20 = 120 buffer channels
        Cs, mid_s = _pad4(Cc), G                                # 116: pitch of a 114-channel group
        self.C, self.Cs, self.Cd = Cc, Cs, Cd

        def blocks(ml):
            return [(_Conv(b.block[0], b.block[1], 'lrelu', dev, cin_map=group_map(bi + 1, nch)),
                     _Conv(b.block[4], b.block[5], 'lrelu', dev)) for bi, b in enumerate(ml)]

        z = lambda c: torch.zeros(N, c, dtype=torch.float32, device=dev)     # noqa: E731
        dense_map = group_map(gf + 1, nch)
        if self.v3:
            heads = [(blocks(net.block0), _Vortex(net.vortex0, dev, dense_map))]
            if self.use_sem:
                heads.append((blocks(net.block2), _Vortex(net.vortex2, dev, dense_map)))
            nh = len(heads)
            tail = _Vortex(net.vortex3, dev, group_map(nh, Cc))
        else:
            heads, nh = [(blocks(net.block), _Vortex(net.vortex, dev, dense_map))], 1
            tail = _Vortex(net.vortex_final, dev, group_map(1, Cc))
        self.two = nh == 2
        self.in_bufs = [z(Cd) for _ in range(nh)]
        cat_stride = nh * Cs
        self.cat = z(cat_stride)
        self.est = torch.empty(1, N, self.P, dtype=torch.float32, device=dev)
        self._keep = [heads, tail]                              # owns every device tensor the plan points at
        self.partial = torch.empty(self.PARTIAL_BLOCKS * 256, dtype=torch.float32, device=dev)
        self.plan = []
        self.side = None
        self.flags = 1 | int(getattr(net, 'conv_flags', 0))        # 1: the pad channels are ours; 64: 1xTF32 (precision 'fast')
        # chains of 1x1 convolutions (the end of every vortex block, the Pred stack) as single launches that keep the
        # activations in tensor memory (csrc/ojdf_conv_chain.cu); OJDF_CHAIN=0: layer by layer
        self.chain = os.environ.get('OJDF_CHAIN', '1') != '0' and bool(getattr(net, 'use_chain', True))

        def conv_step(convs_problems):
            """convs_problems: list of (conv, problem) with identical shapes -> one batched launch."""
            c0 = convs_problems[0][0]
            arr = (ConvProblem * len(convs_problems))(*[p for _, p in convs_problems])
            self.plan.append(('conv', arr, len(convs_problems), c0.cin, c0.cout, c0.taps, c0.act, c0.slope))

        # dense blocks, heads in lock step; block bi reads groups 0..bi and appends group bi+1
        t19 = [z(mid_s) for _ in range(nh)]
        self._keep.append(t19)
        for bi in range(gf):
            conv_step([(heads[hh][0][bi][0], heads[hh][0][bi][0].problem(self.in_bufs[hh], Cd, t19[hh], mid_s)) for hh in range(nh)])
            conv_step([(heads[hh][0][bi][1], heads[hh][0][bi][1].problem(t19[hh], mid_s, self.in_bufs[hh], Cd, (bi + 1) * G))
                       for hh in range(nh)])

        def vortex_steps(vs, srcs, src_stride, dsts):
            """vs: vortex modules run in lock step; srcs[i] -> dsts[i] = (buffer, stride, channel offset)."""
            n = len(vs)
            cin, Cv = vs[0].cin, vs[0].cout
            ps, Cvp = _pad4(cin), _pad4(Cv)
            tb = [[[z(mid_s) for _ in range(2)] for _ in range(4)] for _ in range(n)]
            P1 = [[None, None] + [z(mid_s) for _ in range(2)] for _ in range(n)]   # pool(Y_b), b = 2, 3
            P2 = [z(mid_s) for _ in range(n)]                                       # pool(pool(Y_3))
            chained = self.chain and n <= 2 and Cv <= 128
            br_out = [None if chained else z(4 * Cvp) for _ in range(n)]   # the chain launch never materialises the branch outputs
            self._keep += [tb, P1, P2, br_out]
            # global-pool branch -> bias of the final conv: a long, thin reduction (two launches, one of them a single
            # block) that only the LAST conv of the vortex needs: it runs on a side stream next to the branch convolutions
            self.plan.append(('fork_bias',))
            for i, v in enumerate(vs):
                self.plan.append(('bias', v, srcs[i], src_stride))
            # the first 1x1 convolution of all four branches as ONE launch (the bare products W_b . x, 4 groups of 20 output
            # channels: the activation operand is split once instead of four times); branch 0 gets its bias / BatchNorm /
            # ReLU from an identity "pool", branches 1..3 after their cascaded pools
            Yall = [z(4 * mid_s) for _ in range(n)]
            self._keep.append(Yall)
            conv_step([(vs[i].raw_all, vs[i].raw_all.problem(srcs[i], src_stride, Yall[i], 4 * mid_s)) for i in range(n)])

            def pool_step(items):
                """items: (src, dst, (scale, shift) or None[, identity]) with src = tensor or (tensor, channel offset, stride);
                one launch, ReLU where an epilogue is given."""
                probs = []
                for it in items:
                    a, d, e = it[:3]
                    ident = int(it[3]) if len(it) > 3 else 0
                    ptr, stride = (a[0].data_ptr() + 4 * a[1], a[2]) if isinstance(a, tuple) else (a.data_ptr(), mid_s)
                    probs.append(PoolProblem(ptr, d.data_ptr(), e[0].data_ptr() if e else None, e[1].data_ptr() if e else None,
                                             stride, mid_s, ident))
                arr = (PoolProblem * len(items))(*probs)
                self._keep.append([it[2] for it in items])
                self.plan.append(('pools', arr, len(items), mid_s))

            ysl = lambda i, b: (Yall[i], b * mid_s, 4 * mid_s)   # noqa: E731  (branch b's product inside the merged buffer)
            pool_step([(ysl(i, 0), tb[i][0][0], vs[i].post[0], 1) for i in range(n)] +
                      [(ysl(i, 1), tb[i][1][0], vs[i].post[1]) for i in range(n)] +
                      [(ysl(i, b), P1[i][b], None) for i in range(n) for b in (2, 3)])
            pool_step([(P1[i][2], tb[i][2][0], vs[i].post[2]) for i in range(n)] + [(P1[i][3], P2[i], None) for i in range(n)])
            pool_step([(P2[i], tb[i][3][0], vs[i].post[3]) for i in range(n)])
            conv_step([(vs[i].branches[b][1], vs[i].branches[b][1].problem(tb[i][b][0], mid_s, tb[i][b][1], mid_s))
                       for i in range(n) for b in range(4)])
            conv_step([(vs[i].branches[b][2], vs[i].branches[b][2].problem(tb[i][b][1], mid_s, tb[i][b][0], mid_s))
                       for i in range(n) for b in range(4)])
            if chained:
                # branch-out convs + `final` as one chain launch: per branch 19 -> 114 (+BN+ReLU) stays in tensor memory and is
                # multiplied by its 114-column slice of `final`, the four products add up in the second accumulator
                self.plan.append(('join_bias',))
                inputs = [ChainInput(_ptrs([tb[i][b][0] for i in range(n)]), mid_s, vs[0].branches[b][3].cin) for b in range(4)]
                steps = []
                for b in range(4):
                    c3 = [vs[i].branches[b][3] for i in range(n)]
                    steps.append(ChainStep(_ptrs([c.weights_tc for c in c3]), _ptrs([c.scale for c in c3]), _ptrs([c.shift for c in c3]),
                                           b, c3[0].cin, Cv, 0, 1, 1, c3[0].act, c3[0].slope))
                    fb = [vs[i].final_b[b] for i in range(n)]
                    steps.append(ChainStep(_ptrs([c.weights_tc for c in fb]), _ptrs([c.scale for c in fb]),
                                           _ptrs([vs[i].frame_shift for i in range(n)]), -1, Cv, fb[0].cout, 1, 1 if b == 0 else 0,
                                           2 if b == 3 else 0, fb[0].act, fb[0].slope))
                self.chain_step(inputs, steps, n, [d[0] for d in dsts], [d[2] for d in dsts], dsts[0][1], 1.0)
                return
            conv_step([(vs[i].branches[b][3], vs[i].branches[b][3].problem(tb[i][b][0], mid_s, br_out[i], 4 * Cvp, b * Cvp))
                       for i in range(n) for b in range(4)])
            self.plan.append(('join_bias',))                     # the final conv reads the frame shifts computed on the side stream
            conv_step([(vs[i].final, vs[i].final.problem(br_out[i], 4 * Cvp, dsts[i][0], dsts[i][1], dsts[i][2],
                                                         shift=vs[i].frame_shift)) for i in range(n)])

        vortex_steps([hv[1] for hv in heads], self.in_bufs, Cd, [(self.cat, cat_stride, hh * Cs) for hh in range(nh)])
        vout = z(Cs)
        self._keep.append(vout)
        vortex_steps([tail], [self.cat], cat_stride, [(vout, Cs, 0)])

        pred = []
        for pm in net.pred:
            seq = pm.pred
            if isinstance(seq[5], nn.BatchNorm2d):
                pred += [_Conv(seq[0], seq[1], 'lrelu', dev), _Conv(seq[4], seq[5], 'lrelu', dev)]
            else:                                              # last Pred: conv-BN-LReLU, conv-LReLU, conv-tanh
                pred += [_Conv(seq[0], seq[1], 'lrelu', dev), _Conv(seq[4], None, 'lrelu', dev), _Conv(seq[6], None, 'tanh', dev)]
        pp = [z(Cs) for _ in range(2)]
        self._keep += [pred, pp]
        cur, cs = vout, Cs
        if self.chain and all(c.cout <= 128 for c in pred) and len(pred) <= 12:
            # the whole Pred stack as one chain launch: activations stay in tensor memory between the eleven layers
            # accumulators alternate: the MMAs of layer i+1 start on the K chunks of layer i's output as they are written
            steps = [ChainStep(_ptrs([c.weights_tc]), _ptrs([c.scale]), _ptrs([c.shift]), 0 if i == 0 else -1, c.cin, c.cout, i % 2, 1,
                               2 if i == len(pred) - 1 else 1, c.act, c.slope) for i, c in enumerate(pred)]
            self.chain_step([ChainInput(_ptrs([vout]), Cs, pred[0].cin)], steps, 1, [self.est], [0], self.P, self.scale)
            pred = []
        for i, c in enumerate(pred):
            last = i == len(pred) - 1
            dst, ds = (self.est, self.P) if last else (pp[i % 2], Cs)
            arr = (ConvProblem * 1)(c.problem(cur, cs, dst, ds, 0))
            self.plan.append(('conv', arr, 1, c.cin, c.cout, c.taps, c.act, c.slope, self.scale if last else 1.0))
            cur, cs = dst, ds

    def chain_step(self, inputs, steps, nz, outs, coffs, out_stride, out_mul):
        ia, sa = (ChainInput * len(inputs))(*inputs), (ChainStep * len(steps))(*steps)
        op, oc = (C.c_void_p * nz)(*[o.data_ptr() for o in outs]), (C.c_int * nz)(*[int(c) for c in coffs])
        self.plan.append(('chain', ia, len(inputs), sa, len(steps), nz, op, oc, int(out_stride), float(out_mul)))

    def pack_target(self, frame, sem_frame=None):
        """What Extractor.forward(pack=...) needs to write this engine's input rows itself: (buf_a, buf_b, last_a, last_b, stride)."""
        _lib.require_cuda(frame, sem_frame)
        frame = frame.detach().float().contiguous()
        if self.two:
            assert sem_frame is not None
            sem_frame = sem_frame.detach().float().contiguous()
        return (self.in_bufs[0], self.in_bufs[1] if self.two else None, frame.reshape(-1), sem_frame.reshape(-1) if self.two else None,
                self.Cd)

    def forward(self, vals, wts, frame, sem_frame=None, packed=False):
        """vals / wts: (1,N,P) f32 pixel-major (Extractor output), frame: (1,h,w) f32 depth,
        sem_frame: (1,h,w) f32 normalised labels (1+id)/n_classes (modules/pipeline.py:96).
        Returns est (1,N,P) f32 (a buffer owned by the engine, overwritten by the next call)."""
        _lib.require_cuda(vals, wts, frame, sem_frame)
        dev, N, H, W, P = self.device, self.N, self.h, self.w, self.P
        L = _lib.lib()
        vals = vals.detach().float().contiguous()
        wts = wts.detach().float().contiguous()
        frame = frame.detach().float().contiguous()
        if self.two:
            assert sem_frame is not None
            sem_frame = sem_frame.detach().float().contiguous()
        with torch.cuda.device(dev), _lib.timed('fusionnet', dev):
            main = torch.cuda.current_stream(dev)
            st = main.cuda_stream
            if self.side is None:
                self.side = torch.cuda.Stream(device=dev)
            side = self.side
            if not packed:                                       # packed: the extractor's gather already wrote the input rows
                _lib.check(L.ojdf_pack_fusion_input(vals.data_ptr(), wts.data_ptr(), frame.data_ptr(),
                                                    sem_frame.data_ptr() if self.two else None, N, P, self.in_bufs[0].data_ptr(),
                                                    self.in_bufs[1].data_ptr() if self.two else None, self.Cd, st))
            for step in self.plan:
                kind = step[0]
                if kind == 'conv':
                    _, arr, n, cin, cout, taps, act, slope = step[:8]
                    out_mul = step[8] if len(step) > 8 else 1.0
                    _lib.check(L.ojdf_conv_tc_batched(arr, n, cin, cout, H, W, taps, act, slope, out_mul, 0, self.flags, None, 0, st))
                elif kind == 'chain':
                    _, ia, ni, sa, ns, nz, op, oc, ostride, omul = step
                    _lib.check(L.ojdf_conv_chain(ia, ni, sa, ns, nz, H, W, op, oc, ostride, omul, self.flags, st))
                elif kind == 'pools':
                    _, arr, n, ch = step
                    _lib.check(L.ojdf_avgpool3_batched(arr, n, H, W, ch, 1, st))
                elif kind == 'pool':
                    _, src, ss, ch, dst, ds = step
                    _lib.check(L.ojdf_avgpool3_nhwc(src.data_ptr(), ss, H, W, ch, dst.data_ptr(), ds, st))
                elif kind == 'fork_bias':
                    side.wait_stream(main)                       # the vortex input is complete on the main stream
                elif kind == 'join_bias':
                    main.wait_stream(side)
                else:
                    _, v, src, ss = step
                    _lib.check(L.ojdf_vortex_bias(src.data_ptr(), ss, N, v.cin, v.wg.data_ptr(), v.g_scale.data_ptr(),
                                                  v.g_shift.data_ptr(), v.cout, v.wf1.data_ptr(), v.final.scale.data_ptr(),
                                                  v.final.shift.data_ptr(), v.cout, self.partial.data_ptr(),
                                                  self.PARTIAL_BLOCKS, v.frame_shift.data_ptr(), side.cuda_stream))
        return self.est
