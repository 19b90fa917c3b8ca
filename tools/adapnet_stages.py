#!/usr/bin/env python
"""AdapNet++ stage-2 at 240x320: device time of front (conv1..layer3[0], library), tail engine (libojdf),
SSMA + decoder (library), each replayed as a CUDA graph and timed with CUDA events."""
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
net = AdapNet(fusion_config(240, 320).SEMANTIC_2D_MODEL).to(dev).eval()
net.set_bottleneck_dropout(False)
x1, x2 = torch.randn(1, 3, 240, 320, device=dev), torch.randn(1, 3, 240, 320, device=dev)


def timed(fn, reps=10):
    """Device time of fn() replayed as a CUDA graph (what the pipeline does), so host launch cost is excluded."""
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        out = fn()
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


with torch.no_grad():
    t_front, (f1, f2) = timed(lambda: (net.encoder_mod1.forward_front(x1), net.encoder_mod2.forward_front(x2)))
    t_tail, (e1, e2) = timed(lambda: net._tail([f1[0], f2[0]]))
    t_ssma, (s2, s1, x) = timed(lambda: (net.ssma_s2(f1[1], f2[1]), net.ssma_s1(f1[2], f2[2]), net.ssma_res(e1, e2)))
    t_dec, _ = timed(lambda: net.decoder(x, s1, s2))
    t_all, _ = timed(lambda: net(x1, x2))
    n = net.encoder_mod1.res_n50_enc
    t_stem, y = timed(lambda: n.maxpool(n.relu(n.bn1(n.conv1(x1)))))
    t_l1, y1 = timed(lambda: n.layer1(y))
    t_l2, y2 = timed(lambda: n.layer2(y1))
    t_l30, _ = timed(lambda: n.layer3[0](y2))
    d = net.decoder
    t_dc1, z = timed(lambda: torch.relu(d.deconv1_bn(d.deconv1(x))))
    t_st2, z2 = timed(lambda: d.stage2(d._join(z, s1, d.fuse_conv1)))
    t_st3, _ = timed(lambda: d.stage3(d._join(z2, s2, d.fuse_conv2)))
print('front (2 encoders) %.2f ms | tail engine %.2f | ssma %.2f | decoder %.2f | whole eager %.2f' % (t_front, t_tail, t_ssma, t_dec, t_all))
print('one encoder front: stem %.2f  layer1 %.2f  layer2 %.2f  layer3[0] %.2f' % (t_stem, t_l1, t_l2, t_l30))
print('decoder: deconv1 %.2f  stage2 %.2f  stage3 %.2f' % (t_dc1, t_st2, t_st3))
