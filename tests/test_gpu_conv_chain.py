"""ojdf_conv_chain (csrc/ojdf_conv_chain.cu): chains of 1x1 convolutions kept in tensor memory, against the same layers
in fp64 torch -- FusionNet's Pred stack (modules/model.py:24-52) and the end of a VortexPooling block (four 19 -> 114
conv + BN + ReLU branches concatenated into the 456 -> 114 `final` conv, modules/model.py:131-141,157-159).
Tolerance: max |a-b| <= 5e-5 * max |b| (3xTF32 per layer, errors of up to 11 chained layers)."""
import ctypes as C

import numpy as np
import pytest
import torch

from online_joint_depthfusion_and_semantic_b200 import _lib
from online_joint_depthfusion_and_semantic_b200.modules.fusion_engine import ChainInput, ChainStep, _ptrs

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
ACTS = {0: lambda y: y, 1: lambda y: y.clamp(min=0), 2: lambda y: torch.where(y > 0, y, 0.01 * y), 3: torch.tanh}


def _layer(g, cin, cout):
    """Random 1x1 layer: (weight (cout,cin) f32, scale, shift, packed device image)."""
    L = _lib.lib()
    w = torch.randn(cout, cin, generator=g) / cin ** 0.5
    sc, sh = 0.5 + torch.rand(cout, generator=g), 0.1 * torch.randn(cout, generator=g)
    packed = np.zeros(L.ojdf_conv_tc_weight_floats(cin, cout, 1, 0), np.float32)
    wc = np.ascontiguousarray(w.numpy().reshape(cout, cin, 1))
    _lib.check(L.ojdf_conv_tc_pack_weights(wc.ctypes.data, cin, cout, 1, 0, packed.ctypes.data))
    return w, sc, sh, torch.from_numpy(packed).to(DEV), sc.to(DEV), sh.to(DEV)


def _launch(inputs, steps, nz, H, W, outs, coffs, ostride, out_mul, flags):
    L = _lib.lib()
    ia = (ChainInput * len(inputs))(*inputs)
    sa = (ChainStep * len(steps))(*steps)
    op = (C.c_void_p * nz)(*[o.data_ptr() for o in outs])
    oc = (C.c_int * nz)(*coffs)
    rc = L.ojdf_conv_chain(ia, len(inputs), sa, len(steps), nz, H, W, op, oc, ostride, out_mul, flags, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return rc


@pytest.mark.parametrize('alternate', [0, 1])
@pytest.mark.parametrize('H,W,nz', [(8, 16, 1), (48, 64, 1), (37, 53, 2), (240, 320, 1)])
def test_pred_chain_matches_fp64(H, W, nz, alternate):
    """alternate = 1: consecutive layers use different accumulators (the next layer's MMAs overlap the epilogue chunk by
    chunk); 0: one accumulator (tensor pipe and epilogue warps alternate)."""
    g = torch.Generator().manual_seed(H * 1000 + W + nz)
    widths = [114, 95, 95, 76, 76, 57, 57, 38, 38, 19, 19, 9]
    acts = [2] * 10 + [3]
    N, istr = H * W, 116
    xs = [torch.randn(N, istr, generator=g) for _ in range(nz)]
    layers = [[_layer(g, widths[i], widths[i + 1]) for i in range(11)] for _ in range(nz)]
    refs = []
    for z in range(nz):
        y = xs[z][:, :114].double()
        for i, (w, sc, sh, *_r) in enumerate(layers[z]):
            y = ACTS[acts[i]](y @ w.double().t() * sc.double() + sh.double())
        refs.append(0.25 * y)
    xd = [x.to(DEV) for x in xs]
    outs = [torch.full((N, 9), 7.0, device=DEV) for _ in range(nz)]
    inputs = [ChainInput(_ptrs(xd), istr, 114)]
    steps = []
    for i in range(11):
        steps.append(ChainStep(_ptrs([layers[z][i][3] for z in range(nz)]), _ptrs([layers[z][i][4] for z in range(nz)]),
                               _ptrs([layers[z][i][5] for z in range(nz)]), 0 if i == 0 else -1, widths[i], widths[i + 1],
                               (i % 2) * alternate, 1, 2 if i == 10 else 1, acts[i], 0.01))
    assert _launch(inputs, steps, nz, H, W, outs, [0] * nz, 9, 0.25, 0) == 0
    for y, o in zip(refs, outs):
        assert float((o.cpu().double() - y).abs().max()) <= 5e-5 * float(y.abs().max())


@pytest.mark.parametrize('H,W,nz', [(48, 64, 2), (240, 320, 2)])
def test_vortex_tail_chain_matches_fp64(H, W, nz):
    """out = scale_f * sum_b F_b . relu(scale_b * (W_b . h_b) + shift_b) + shift_f, written at channel offset z * 116 of a
    232-wide buffer (the two heads side by side) through the TMA-store epilogue (pad channels owned: flag 1)."""
    g = torch.Generator().manual_seed(H + W)
    N, mid, Cv = H * W, 19, 114
    hs = [[torch.randn(N, 20, generator=g) for _ in range(4)] for _ in range(nz)]
    br = [[_layer(g, mid, Cv) for _ in range(4)] for _ in range(nz)]
    fin = [[_layer(g, Cv, Cv) for _ in range(4)] for _ in range(nz)]           # the four 114-column slices of `final`
    fsc = [0.5 + torch.rand(Cv, generator=g) for _ in range(nz)]
    fsh = [0.1 * torch.randn(Cv, generator=g) for _ in range(nz)]
    refs = []
    for z in range(nz):
        acc = torch.zeros(N, Cv, dtype=torch.float64)
        for b in range(4):
            w, sc, sh = br[z][b][:3]
            t = (hs[z][b][:, :mid].double() @ w.double().t() * sc.double() + sh.double()).clamp(min=0)
            acc += t @ fin[z][b][0].double().t()
        refs.append(acc * fsc[z].double() + fsh[z].double())
    hd = [[h.to(DEV) for h in hz] for hz in hs]
    fscd, fshd = [t.to(DEV) for t in fsc], [t.to(DEV) for t in fsh]
    out = torch.full((N, 232), 7.0, device=DEV)
    inputs = [ChainInput(_ptrs([hd[z][b] for z in range(nz)]), 20, mid) for b in range(4)]
    steps = []
    for b in range(4):
        steps.append(ChainStep(_ptrs([br[z][b][3] for z in range(nz)]), _ptrs([br[z][b][4] for z in range(nz)]),
                               _ptrs([br[z][b][5] for z in range(nz)]), b, mid, Cv, 0, 1, 1, 1, 0.0))
        steps.append(ChainStep(_ptrs([fin[z][b][3] for z in range(nz)]), _ptrs(fscd), _ptrs(fshd), -1, Cv, Cv, 1, 1 if b == 0 else 0,
                               2 if b == 3 else 0, 0, 0.0))
    assert _launch(inputs, steps, nz, H, W, [out] * nz, [116 * z for z in range(nz)], 232, 1.0, 1) == 0
    o = out.cpu().double()
    for z in range(nz):
        got = o[:, 116 * z:116 * z + Cv]
        assert float((got - refs[z]).abs().max()) <= 5e-5 * float(refs[z].abs().max())
    if nz == 1:
        assert bool((o[:, 116:] == 7.0).all())


def test_chain_rejects_bad_arguments():
    x = torch.zeros(128, 32, device=DEV)
    inputs = [ChainInput(_ptrs([x]), 32, 32)]
    ok = ChainStep(_ptrs([x]), _ptrs([x]), _ptrs([x]), 0, 32, 32, 0, 1, 2, 0, 0.0)
    out = torch.zeros(128, 32, device=DEV)
    assert _launch(inputs, [ok], 1, 8, 16, [out], [0], 32, 1.0, 0) == 0
    bad_src = ChainStep(_ptrs([x]), _ptrs([x]), _ptrs([x]), -1, 32, 32, 0, 1, 2, 0, 0.0)        # nothing in tensor memory yet
    assert _launch(inputs, [bad_src], 1, 8, 16, [out], [0], 32, 1.0, 0) == -1
    wide = ChainStep(_ptrs([x]), _ptrs([x]), _ptrs([x]), 0, 32, 160, 0, 1, 2, 0, 0.0)          # more than 128 output channels
    assert _launch(inputs, [wide], 1, 8, 16, [out], [0], 32, 1.0, 0) == -1
    no_out = ChainStep(_ptrs([x]), _ptrs([x]), _ptrs([x]), 0, 32, 32, 0, 1, 1, 0, 0.0)         # the last step must be the output step
    assert _launch(inputs, [no_out], 1, 8, 16, [out], [0], 32, 1.0, 0) == -1
