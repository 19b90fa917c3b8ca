"""Host-side logic of the network engines that needs no GPU: the phase decomposition of the decoder's transposed
convolutions, the padded channel-group maps with their zero weight rows, and the pool / 1x1-conv commutation the
FusionNet engine relies on (modules/adapnet_engine.py, modules/fusion_engine.py)."""
import numpy as np
import pytest
import torch
from torch.nn import functional as F

from online_joint_depthfusion_and_semantic_b200.modules.adapnet_engine import deconv_phase_weights
from online_joint_depthfusion_and_semantic_b200.modules.fusion_engine import _Conv, group_map


@pytest.mark.parametrize('k,s,p,cin,cout', [(4, 2, 1, 5, 7), (8, 4, 2, 6, 3)])
def test_transposed_conv_equals_its_phase_convolutions(k, s, p, cin, cout):
    """Decoder.deconv1 / stage2[6] (k=4, s=2, p=1) and stage3[8] (k=8, s=4, p=2), modules/adapnet.py:138-148."""
    g = torch.Generator().manual_seed(k)
    x = torch.randn(1, cin, 9, 11, generator=g, dtype=torch.float64)
    w = torch.randn(cin, cout, k, k, generator=g, dtype=torch.float64)
    ref = F.conv_transpose2d(x, w, stride=s, padding=p)
    out = torch.zeros_like(ref)
    phases = deconv_phase_weights(w, s, p)
    assert len(phases) == s * s
    for a, b, wp, mask in phases:
        assert bin(mask).count('1') == 4                       # two live taps per axis
        for t in range(9):
            if not (mask >> t) & 1:
                assert float(wp[:, :, t // 3, t % 3].abs().max()) == 0.0
        out[:, :, a::s, b::s] = F.conv2d(x, wp, padding=1)
    assert float((out - ref).abs().max()) < 1e-12


def test_group_map_and_zero_weight_rows():
    """Dense-block buffer of FusionNet: 19-channel groups at pitch 20 (fusion_engine.py); the consumer's expanded
    weights must reproduce conv + BatchNorm of the unpadded tensor whatever the pad channels hold."""
    pos, width = group_map(3, 19)
    assert width == 60 and pos[:3] == [0, 1, 2] and pos[19] == 20 and pos[38] == 40 and len(pos) == 57
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(57, 19, 1)
    bn = torch.nn.BatchNorm2d(19).eval()
    bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_()
    c = _Conv(conv, bn, 'none', 'cpu', cin_map=(pos, width))
    assert c.cin == 60 and c.cout == 19 and c.taps == 1
    x = torch.randn(1, 57, 6, 5)
    ref = bn(conv(x)).detach()[0].permute(1, 2, 0).reshape(30, 19)
    xp = torch.full((30, 60), 1234.5)                          # garbage in the pad channels
    xp[:, pos] = x[0].permute(1, 2, 0).reshape(30, 57)
    wmat = _unpack_tc(c.weights_tc, 60, 19, 1)[0]              # (cout, cin) = hi + lo of the packed images, tap 0
    got = (xp.double() @ wmat.t().double()).float() * c.scale + c.shift
    assert float((got - ref).abs().max()) < 1e-4 * float(ref.abs().max())
    assert float(wmat[:, [19, 39, 59]].abs().max()) == 0.0     # zero weights at the pad channels


def _unpack_tc(packed, cin, cout, taps):
    """Inverse of ojdf_conv_tc_pack_weights (include/ojdf.h): [group][tap][K chunk of 32][hi | lo][npad rows][32 floats],
    16-byte chunk c of row r stored at chunk c ^ (r & 7); returns (taps, cout, cin) = hi + lo."""
    import ctypes as C
    from online_joint_depthfusion_and_semantic_b200 import _lib
    npad, groups = C.c_int(), C.c_int()
    _lib.check(_lib.lib().ojdf_conv_tc_layout(cout, 0, C.byref(npad), C.byref(groups)))
    npad, groups = npad.value, groups.value
    nkc = (cin + 31) // 32
    img = packed.reshape(groups, taps, nkc, 2, npad, 32)
    out = torch.zeros(taps, cout, cin)
    for g in range(groups):
        for r in range(npad):
            co = g * npad + r
            if co >= cout:
                continue
            for kc in range(nkc):
                for k in range(32):
                    ci = kc * 32 + k
                    if ci >= cin:
                        continue
                    col = ((k >> 2) ^ (r & 7)) * 4 + (k & 3)
                    out[:, co, ci] = img[g, :, kc, 0, r, col] + img[g, :, kc, 1, r, col]
    return out


def test_average_pools_commute_with_the_1x1_convolution():
    """VortexPooling (modules/model.py:114-135): W.pool^b(x) == pool^b(W.x) for AvgPool2d(3, 1, 1) with the zero
    padding counted in the divisor -- the identity behind pooling 19 channels instead of 114."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 12, 10, 13, generator=g, dtype=torch.float64)
    w = torch.randn(5, 12, 1, 1, generator=g, dtype=torch.float64)
    pool = lambda t: F.avg_pool2d(t, 3, stride=1, padding=1, count_include_pad=True)     # noqa: E731
    a, b = x, F.conv2d(x, w)
    for _ in range(3):
        a, b = pool(a), pool(b)
        assert float((F.conv2d(a, w) - b).abs().max()) < 1e-13


def test_merged_vortex_first_layer_and_final_slices():
    """FusionNetEngine's host transforms of a VortexPooling block (modules/model.py:100-161): the merged first-layer image
    holds branch b's 1x1 weights in output rows [20 b, 20 b + 19) (pad rows zero, input columns at their padded positions),
    and the four `final` slices of the chain launch are the columns of the concatenated convolution in branch order."""
    from online_joint_depthfusion_and_semantic_b200.modules.fusion_engine import _Vortex
    from online_joint_depthfusion_and_semantic_b200.modules.model import VortexPooling
    torch.manual_seed(3)
    cin, mid, cout = 38, 19, 114
    m = VortexPooling(cin, mid, cout, (6, 5)).eval()
    pos, width = group_map(2, 19)                              # two 19-channel groups at pitch 20
    v = _Vortex(m, 'cpu', (pos, width))
    w_all = _unpack_tc(v.raw_all.weights_tc, width, 4 * 20, 1)[0]            # (80, 40)
    for b, br in enumerate(m.branches):
        wb = br[0].weight.detach().reshape(mid, cin)
        assert torch.equal(w_all[20 * b:20 * b + mid][:, pos], wb)
        assert float(w_all[20 * b + mid].abs().max()) == 0.0                  # pad row of the group
    assert float(w_all[:, [19, 39]].abs().max()) == 0.0                       # pad input channels
    wf = m.final[0].weight.detach().reshape(cout, 5 * cout)
    for b in range(4):
        fb = _unpack_tc(v.final_b[b].weights_tc, cout, cout, 1)[0]
        assert torch.equal(fb, wf[:, cout * (1 + b):cout * (2 + b)])
    # branch 0's deferred epilogue = its folded BatchNorm
    bn = m.branches[0][1]
    s = bn.weight.detach().double() / torch.sqrt(bn.running_var.double() + bn.eps)
    assert torch.allclose(v.post[0][0][:mid].double(), s, rtol=1e-6)
