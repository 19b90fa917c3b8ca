#!/usr/bin/env python
"""AdapNet++ stage-2 forward timing at 240x320 under different library settings (experiment)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from online_joint_depthfusion_and_semantic_b200.config import fusion_config  # noqa: E402
from online_joint_depthfusion_and_semantic_b200.cuda_graph import GraphedCall  # noqa: E402
from online_joint_depthfusion_and_semantic_b200.modules.adapnet import AdapNet  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device('cuda:0')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    cfg = fusion_config(240, 320)
    net = AdapNet(cfg.SEMANTIC_2D_MODEL).to(dev).eval()
    net.set_bottleneck_dropout(False)
    x1, x2 = torch.randn(1, 3, 240, 320, device=dev), torch.randn(1, 3, 240, 320, device=dev)
    with torch.no_grad():
        ref = net(x1, x2)[0].clone()
        print('eager                      %.2f ms' % timeit(lambda: net(x1, x2)))
        g = GraphedCall(lambda a, b: net(a, b)[0])
        print('graph                      %.2f ms' % timeit(lambda: g(x1, x2)))
        torch.backends.cudnn.benchmark = True
        g2 = GraphedCall(lambda a, b: net(a, b)[0])
        out = g2(x1, x2)
        print('graph + cudnn.benchmark    %.2f ms   max rel diff %.2e' % (timeit(lambda: g2(x1, x2)), float((out - ref).abs().max() / ref.abs().max())))
        if hasattr(net, 'set_folded'):
            net.set_folded(True)
            g3 = GraphedCall(lambda a, b: net(a, b)[0])
            out = g3(x1, x2)
            print('graph + benchmark + folded %.2f ms   max rel diff %.2e' % (timeit(lambda: g3(x1, x2)), float((out - ref).abs().max() / ref.abs().max())))
            net.set_folded(False)
        netc = net.to(memory_format=torch.channels_last)
        y1, y2 = x1.contiguous(memory_format=torch.channels_last), x2.contiguous(memory_format=torch.channels_last)
        g4 = GraphedCall(lambda a, b: netc(a, b)[0])
        out = g4(y1, y2)
        print('graph + benchmark + NHWC   %.2f ms   max rel diff %.2e' % (timeit(lambda: g4(y1, y2)), float((out - ref).abs().max() / ref.abs().max())))
        torch.backends.cudnn.allow_tf32 = True
        g5 = GraphedCall(lambda a, b: netc(a, b)[0])
        out = g5(y1, y2)
        print('(TF32, NHWC, for scale)    %.2f ms   max rel diff %.2e' % (timeit(lambda: g5(y1, y2)), float((out - ref).abs().max() / ref.abs().max())))


if __name__ == '__main__':
    main()
