"""The tcgen05 tap GEMM (csrc/ojdf_conv_tc.cu) through the C ABI against an fp64 torch convolution of the
same layer (conv + per-channel scale/shift + residual + activation), i.e. the reference's
Conv2d + BatchNorm2d(eval) + activation of modules/model.py:4-52 and modules/adapnet.py:12-84.
Tolerance: max |a-b| <= 5e-5 * max |b| per layer (3xTF32 keeps ~1e-6 for short reductions; the 4608-term
reduction of a 512-channel 3x3 reaches 3.5e-5), well inside the 1e-4 the north star allows end to end."""
import ctypes as C

import numpy as np
import pytest
import torch

from online_joint_depthfusion_and_semantic_b200 import _lib
from online_joint_depthfusion_and_semantic_b200.modules.fusion_engine import ConvProblem

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')

# (name, H, W, cin, cout, taps, dil, in_stride, out_stride, out_coff, act, residual, n_problems, npad_req, flags)
CASES = [
    ('1x1 one tile one chunk', 8, 16, 32, 32, 1, 1, 32, 32, 0, 0, False, 1, 0, 0),
    ('1x1 cin 19 of stride 116, ragged width', 8, 16, 19, 19, 1, 1, 116, 20, 0, 2, False, 1, 0, 0),
    ('1x1 114 -> 114 unaligned width, plain stores', 16, 32, 114, 114, 1, 1, 116, 116, 0, 1, False, 1, 0, 0),
    ('1x1 114 -> 114 pad channels owned (TMA store)', 16, 32, 114, 114, 1, 1, 116, 116, 0, 1, False, 1, 0, 1),
    ('3x3 one tile', 8, 16, 32, 32, 9, 1, 32, 32, 0, 0, False, 1, 0, 0),
    ('3x3 dense block into an unaligned channel offset', 48, 64, 95, 19, 9, 1, 116, 116, 95, 2, False, 2, 0, 0),
    ('3x3 dense block into a 4-aligned group', 48, 64, 100, 19, 9, 1, 120, 120, 100, 2, False, 2, 0, 1),
    ('3x3 mixed dilations 3/2, ragged image (per-tap boxes)', 37, 53, 19, 19, 9, 3, 20, 20, 0, 1, False, 4, 0, 0),
    ('3x3 dilation 27 (taps mostly outside)', 48, 64, 19, 19, 9, 27, 20, 20, 0, 1, False, 8, 0, 0),
    ('3x3 dilation 2 halo box', 30, 40, 64, 64, 9, 2, 64, 64, 0, 1, False, 1, 0, 0),
    ('1x1 456 -> 114', 48, 64, 456, 114, 1, 1, 456, 116, 0, 0, False, 2, 0, 1),
    ('1x1 -> 9 tanh, row stride 9', 48, 64, 19, 9, 1, 1, 116, 9, 0, 3, False, 1, 0, 0),
    ('1x1 256 -> 256 (2 groups) residual sigmoid', 30, 40, 256, 256, 1, 1, 256, 256, 0, 4, True, 1, 0, 0),
    ('3x3 64 -> 64 residual relu, two problems', 60, 80, 64, 64, 9, 1, 64, 64, 0, 1, True, 2, 0, 0),
    ('3x3 15x20 512 -> 512 (4 groups)', 15, 20, 512, 512, 9, 1, 512, 512, 0, 1, False, 1, 0, 0),
    ('1x1 15x20 1024 -> 256 narrow groups of 32', 15, 20, 1024, 256, 1, 1, 1024, 512, 0, 1, False, 2, 32, 0),
    ('3x3 240x320 dense block, many groups per CTA', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 0, 1),
    ('1x1 one M-tile per group (flag 2)', 48, 64, 95, 19, 1, 1, 116, 116, 95, 2, False, 2, 0, 2),
    ('3x3 15x20 512 -> 256 mixed dilations, residual, split K', 15, 20, 512, 256, 9, 4, 512, 256, 0, 1, True, 4, 0, 0),
    ('3x3 15x20 512 -> 512 without the K split (flag 4096)', 15, 20, 512, 512, 9, 1, 512, 512, 0, 1, False, 1, 0, 4096),
    ('3x3 15x20 512 -> 256 split K finished inside the kernel (flag 16384)', 15, 20, 512, 256, 9, 4, 512, 256, 0, 1, True, 4, 0, 16384),
    ('1x1 15x20 1024 -> 256 split K inside the kernel, sigmoid gate (act 5)', 15, 20, 1024, 256, 1, 1, 1024, 256, 0, 5, True, 2, 0, 16384),
    ('3x3 30x40 48 -> 48 sigmoid gate (act 5) on the shared-memory-operand kernel', 30, 40, 48, 48, 9, 1, 48, 48, 0, 5, True, 1, 0, 0),
    # the swapped-operand kernel (csrc/ojdf_conv_wt.cu): pixel tiles, 64-channel groups, strided output offset, and its
    # cluster / distributed-shared-memory split-K experiment (flag 524288)
    ('wt: 1x1 30x40 128 -> 512 residual, two problems', 30, 40, 128, 512, 1, 1, 128, 512, 0, 1, True, 2, 0, 0),
    ('wt: 3x3 30x40 280 -> 256 five pixel tiles', 30, 40, 280, 256, 9, 1, 280, 256, 0, 1, False, 1, 0, 0),
    ('wt: 3x3 60x80 64 -> 64 (64-channel group), dilation 2', 60, 80, 64, 64, 9, 2, 64, 64, 0, 1, False, 2, 0, 0),
    ('wt: 1x1 15x20 1024 -> 256 cluster split K (flag 524288)', 15, 20, 1024, 256, 1, 1, 1024, 512, 0, 1, True, 2, 0, 524288),
    ('wt: 3x3 23x37 96 -> 128 ragged map, cluster split K, sigmoid gate', 23, 37, 96, 128, 9, 3, 96, 128, 0, 5, True, 2, 0, 524288),
]


def _run(case):
    name, H, W, cin, cout, taps, dil, istr, ostr, ocoff, act, use_res, nprob, npad_req, flags = case
    L = _lib.lib()
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    k = 3 if taps == 9 else 1
    keep, probs, refs, outs = [], [], [], []
    for i in range(nprob):
        x = torch.randn(H * W, istr, generator=g)                           # channels >= cin are garbage on purpose
        w = torch.randn(cout, cin, k, k, generator=g) / (cin * taps) ** 0.5
        sc, sh = 0.5 + torch.rand(cout, generator=g), 0.1 * torch.randn(cout, generator=g)
        res = torch.randn(H * W, cout, generator=g) if use_res else None
        packed = np.zeros(L.ojdf_conv_tc_weight_floats(cin, cout, taps, npad_req), np.float32)
        wc = np.ascontiguousarray(w.numpy().reshape(cout, cin, taps))
        _lib.check(L.ojdf_conv_tc_pack_weights(wc.ctypes.data, cin, cout, taps, npad_req, packed.ctypes.data))
        d = (dil if nprob == 1 else max(1, dil - i % 2)) if taps == 9 else 1
        xin = x[:, :cin].double().t().reshape(1, cin, H, W)
        y = torch.nn.functional.conv2d(xin, w.double(), padding=d * (k // 2), dilation=d)[0].reshape(cout, H * W).t()
        y = y * sc.double() + sh.double()
        if act == 5:                                                        # SSMA gate: sigmoid(conv) * gated tensor
            y = torch.sigmoid(y) * res.double()
        else:
            if use_res:
                y = y + res.double()
            y = {0: y, 1: y.clamp(min=0), 2: torch.where(y > 0, y, 0.01 * y), 3: torch.tanh(y), 4: torch.sigmoid(y)}[act]
        out = torch.full((H * W, ostr), 7.0, device=DEV)
        t = [x.to(DEV), torch.from_numpy(packed).to(DEV), sc.to(DEV), sh.to(DEV), out, res.to(DEV) if use_res else None]
        keep.append(t)
        probs.append(ConvProblem(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), out.data_ptr(),
                                 t[5].data_ptr() if use_res else None, istr, ostr, ocoff, d, cout if use_res else 0, 0, 0, 0, 0, 0))
        refs.append(y)
        outs.append(out)
    arr = (ConvProblem * nprob)(*probs)
    scratch = torch.zeros((32 << 20) // 4, dtype=torch.float32, device=DEV)     # lets small maps split their K loop (zero: slice counters)
    _lib.check(L.ojdf_conv_tc_batched(arr, nprob, cin, cout, H, W, taps, act, 0.01, 1.0, npad_req, flags, scratch.data_ptr() if scratch is not None else None,
                                      scratch.numel() * 4 if scratch is not None else 0, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return refs, [o.cpu().double() for o in outs]


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_conv_tc_matches_fp64(case):
    _, H, W, cin, cout, taps, dil, istr, ostr, ocoff, act, use_res, nprob, npad_req, flags = case
    refs, outs = _run(case)
    pad = (-cout) % 4 if flags & 1 else 0
    for y, o in zip(refs, outs):
        got = o[:, ocoff:ocoff + cout]
        assert float((got - y).abs().max()) <= 5e-5 * float(y.abs().max())
        # nothing outside the layer's own channels is touched, except the pad channels the caller declared its own:
        # those receive zeros (TMA-store epilogue) or stay as they were (split-K / plain-store epilogues)
        assert bool((o[:, :ocoff] == 7.0).all()) and bool((o[:, ocoff + cout + pad:] == 7.0).all())
        if pad:
            p_ = o[:, ocoff + cout:ocoff + cout + pad]
            assert bool((p_ == 0.0).all()) or bool((p_ == 7.0).all())


def test_conv_tc_rejects_bad_arguments():
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    x = torch.zeros(128, 32, device=DEV)
    p = ConvProblem(x.data_ptr() + 4, x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), None, 32, 32, 0, 1, 0, 0, 0, 0, 0, 0)
    arr = (ConvProblem * 1)(p)
    assert L.ojdf_conv_tc_batched(arr, 1, 32, 32, 8, 16, 1, 0, 0.0, 1.0, 0, 0, None, 0, st) == -1       # misaligned input
    assert L.ojdf_conv_tc_batched(arr, 1, 32, 32, 8, 16, 4, 0, 0.0, 1.0, 0, 0, None, 0, st) == -1       # taps not 1 or 9
    assert L.ojdf_conv_tc_batched(None, 1, 32, 32, 8, 16, 1, 0, 0.0, 1.0, 0, 0, None, 0, st) == -1


@pytest.mark.parametrize('cout', [80, 128])
def test_conv_tc_strided_input_matches_stride2_conv(cout):
    """in_step = 2: a 1x1 convolution with stride 2 (ResNet down-sampling, torchvision Bottleneck.downsample) and
    a stride-2 3x3 expressed as the stride-1 3x3 followed by a strided 1x1 read (modules/adapnet_engine.py)."""
    L = _lib.lib()
    g = torch.Generator().manual_seed(9)
    Hin, Win, cin = 30, 40, 96                  # cout 80: pixel-major kernel; 128: the swapped-operand kernel (ojdf_conv_wt.cu)
    Ho, Wo = Hin // 2, Win // 2
    x = torch.randn(Hin * Win, cin, generator=g)
    w = torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5
    sc, sh = 0.5 + torch.rand(cout, generator=g), 0.1 * torch.randn(cout, generator=g)
    ref = torch.nn.functional.conv2d(x.double().t().reshape(1, cin, Hin, Win), w.double(), stride=2)[0]
    ref = (ref.reshape(cout, Ho * Wo).t() * sc.double() + sh.double()).clamp(min=0)
    packed = np.zeros(L.ojdf_conv_tc_weight_floats(cin, cout, 1, 0), np.float32)
    wc = np.ascontiguousarray(w.numpy().reshape(cout, cin, 1))
    _lib.check(L.ojdf_conv_tc_pack_weights(wc.ctypes.data, cin, cout, 1, 0, packed.ctypes.data))
    xd, pd, scd, shd = x.to(DEV), torch.from_numpy(packed).to(DEV), sc.to(DEV), sh.to(DEV)
    out = torch.full((Ho * Wo, cout), 7.0, device=DEV)
    p = ConvProblem(xd.data_ptr(), pd.data_ptr(), scd.data_ptr(), shd.data_ptr(), out.data_ptr(), None, cin, cout, 0, 1, 0, 2, Win, 0, 0, 0)
    arr = (ConvProblem * 1)(p)
    _lib.check(L.ojdf_conv_tc_batched(arr, 1, cin, cout, Ho, Wo, 1, 1, 0.0, 1.0, 0, 0, None, 0, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert float((out.cpu().double() - ref).abs().max()) <= 5e-5 * float(ref.abs().max())
    bad = ConvProblem(xd.data_ptr(), pd.data_ptr(), scd.data_ptr(), shd.data_ptr(), out.data_ptr(), None, cin, cout, 0, 1, 0, 2, Wo, 0, 0, 0)
    assert L.ojdf_conv_tc_batched((ConvProblem * 1)(bad), 1, cin, cout, Ho, Wo, 1, 1, 0.0, 1.0, 0, 0, None, 0,
                                  torch.cuda.current_stream().cuda_stream) == -1          # input row narrower than the reads
