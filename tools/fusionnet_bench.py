#!/usr/bin/env python
"""FusionNet_v3 forward on libojdf's conv kernels: timing (+ optional ncu range)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from online_joint_depthfusion_and_semantic_b200 import _lib  # noqa: E402
from online_joint_depthfusion_and_semantic_b200.config import fusion_config  # noqa: E402
from online_joint_depthfusion_and_semantic_b200.modules.model import FusionNet_v3  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--h', type=int, default=240)
    ap.add_argument('--w', type=int, default=320)
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--profile', action='store_true')
    ap.add_argument('--torch', action='store_true', help='also time the cuDNN fp32 forward of the same module')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    cfg = fusion_config(a.h, a.w)
    cfg.FUSION_MODEL.resx, cfg.FUSION_MODEL.resy = a.w, a.h
    net = FusionNet_v3(cfg.FUSION_MODEL).to(dev).eval()
    N = a.h * a.w
    vals, wts = 0.05 * torch.randn(1, N, 9, device=dev), torch.rand(1, N, 9, device=dev)
    frame, sem = torch.rand(1, a.h, a.w, device=dev) * 2, torch.rand(1, a.h, a.w, device=dev)
    with torch.no_grad():
        for _ in range(3):
            net.forward_pixel_major(vals, wts, frame, sem)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        if a.profile:
            torch.cuda.profiler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            net.forward_pixel_major(vals, wts, frame, sem)
        e1.record()
        torch.cuda.synchronize()
        if a.profile:
            torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1) / a.reps
        flop = 78.15e9 * (a.h * a.w) / 76800.0
        print('engine forward %.3f ms  %.2f TFLOP/s  (%d launches/forward)' % (ms, flop / ms / 1e9, (_lib.launch_count() - l0) // a.reps))
        if a.torch:
            x = {'tsdf_values': vals.view(1, a.h, a.w, 9).permute(0, 3, 1, 2).contiguous(),
                 'tsdf_weights': wts.view(1, a.h, a.w, 9).permute(0, 3, 1, 2).contiguous(),
                 'tsdf_frame': frame[:, None], 'semantic_frame': sem[:, None]}
            net.use_engine = False
            for _ in range(2):
                net(x)
            e0.record()
            for _ in range(a.reps):
                net(x)
            e1.record()
            torch.cuda.synchronize()
            print('torch/cuDNN fp32 forward %.3f ms' % (e0.elapsed_time(e1) / a.reps))


if __name__ == '__main__':
    main()
