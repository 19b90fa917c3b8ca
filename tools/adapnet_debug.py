#!/usr/bin/env python
"""Stage-by-stage comparison of the whole-network AdapNet++ engine with the torch forward (bring-up aid)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from online_joint_depthfusion_and_semantic_b200.config import fusion_config  # noqa: E402
from online_joint_depthfusion_and_semantic_b200.modules.adapnet import AdapNet  # noqa: E402

dev = torch.device('cuda:0')
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (240, 320)
net = AdapNet(fusion_config(h, w).SEMANTIC_2D_MODEL).to(dev).eval()
net.set_bottleneck_dropout(False)
x1, x2 = torch.randn(1, 3, h, w, device=dev), torch.randn(1, 3, h, w, device=dev)


def nhwc(t):
    return t[0].permute(1, 2, 0).reshape(-1, t.shape[1])


def cmp(name, got, ref):
    err = float((got - ref).abs().max())
    print('%-28s max err %.3e  scale %.3e  %s' % (name, err, float(ref.abs().max()), 'ok' if err <= 1e-4 * float(ref.abs().max()) else 'BAD'))


with torch.no_grad():
    out = net(x1, x2)
    e = net._full_engine
    torch.cuda.synchronize()
    encs = [net.encoder_mod1, net.encoder_mod2]
    pres, s2s, s1s = [], [], []
    for i, (enc, x) in enumerate(zip(encs, (x1, x2))):
        n = enc.res_n50_enc
        y = n.maxpool(n.relu(n.bn1(n.conv1(x))))
        cmp('enc%d stem' % i, e.S0[i], nhwc(y))
        y = n.layer1(y)
        cmp('enc%d layer1' % i, e.L1[i], nhwc(y))
        s2 = enc.enc_skip2_conv_bn(enc.enc_skip2_conv(y))
        cmp('enc%d skip2' % i, e.SK2[:, 24 * i:24 * i + 24], nhwc(s2))
        y = n.layer2(y)
        cmp('enc%d layer2' % i, e.L2[i], nhwc(y))
        s1 = enc.enc_skip1_conv_bn(enc.enc_skip1_conv(y))
        cmp('enc%d skip1' % i, e.SK1[:, 24 * i:24 * i + 24], nhwc(s1))
        y = n.layer3[0](y)
        cmp('enc%d layer3[0]' % i, e.tail.X[i][0][:, :1024], nhwc(y))
        for u in list(n.layer3)[1:]:
            y = u(y)
        y = n.layer4(y)
        pres.append(y); s2s.append(s2); s1s.append(s1)
    a1, a2 = net.eASPP_mod1(pres[0]), net.eASPP_mod2(pres[1])
    cmp('eASPP 1', e.FX[:, :256], nhwc(a1))
    cmp('eASPP 2', e.FX[:, 256:], nhwc(a2))
    k2, k1, xr = net.ssma_s2(s2s[0], s2s[1]), net.ssma_s1(s1s[0], s1s[1]), net.ssma_res(a1, a2)
    cmp('ssma skip2', e.skip2, nhwc(k2))
    cmp('ssma skip1', e.skip1, nhwc(k1))
    cmp('ssma res', e.X16, nhwc(xr))
    d = net.decoder
    x = torch.relu(d.deconv1_bn(d.deconv1(xr)))
    j1 = d._join(x, k1, d.fuse_conv1)
    cmp('join1', e.J1, nhwc(j1))
    x = d.stage2[:6](j1)
    cmp('stage2 convs', e.U2, nhwc(x))
    x = d.stage2[6:](x)
    j2 = d._join(x, k2, d.fuse_conv2)
    cmp('join2', e.J2, nhwc(j2))
    y3 = d.stage3[:8](j2)
    cmp('stage3 convs', e.V3[:, :30], nhwc(y3))
    ref = d.stage3[8:](y3)
    cmp('logits', out[0], ref)
