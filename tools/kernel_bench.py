#!/usr/bin/env python
"""Kernel-only timing of the extract / integrate path on synthetic frames (no networks).

    python tools/kernel_bench.py [--h 240 --w 320 --grid 256 --frames 12 --reps 3]

Each frame runs Extractor.forward + Integrator.forward (FrameUpdate form, semantics on) with a
seeded pseudo network output.  Between frames a 512 MB buffer is written so the voxel volumes and
the workspace never sit in L2 (the real pipeline has >1 GB of conv activations between the two).
Prints CUDA-event times per C-ABI call; wrap in ncu for per-kernel detail."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from online_joint_depthfusion_and_semantic_b200 import _lib  # noqa: E402
from online_joint_depthfusion_and_semantic_b200.config import fusion_config  # noqa: E402
from online_joint_depthfusion_and_semantic_b200.modules import Extractor, Integrator  # noqa: E402
from online_joint_depthfusion_and_semantic_b200.modules.integrator import FrameUpdate  # noqa: E402
from online_joint_depthfusion_and_semantic_b200.synthetic import SyntheticScene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--h', type=int, default=240)
    ap.add_argument('--w', type=int, default=320)
    ap.add_argument('--grid', type=int, default=256)
    ap.add_argument('--frames', type=int, default=12)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--no-flush', action='store_true')
    ap.add_argument('--profile', action='store_true', help='cudaProfilerStart/Stop around the last rep')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    h, w, G = a.h, a.w, a.grid
    N = h * w
    scene = SyntheticScene(grid=G, h=h, w=w, n_frames=a.frames, seed=1)
    cfg = fusion_config(h, w, semantic_strategy='gt')
    ex, integ = Extractor(cfg), Integrator(cfg)
    tsdf = torch.full((G, G, G), 0.1, dtype=torch.float16, device=dev)
    wvol = torch.zeros((G, G, G), dtype=torch.float16, device=dev)
    ids = torch.zeros((G, G, G), dtype=torch.uint8, device=dev)
    sc = torch.zeros((G, G, G), dtype=torch.float16, device=dev)
    frames = [scene.frame(i, device=dev) for i in range(a.frames)]
    for b in frames:
        b['extrinsics'], b['intrinsics'] = b['extrinsics'].cpu(), b['intrinsics'].cpu()
    g = torch.Generator(device='cpu').manual_seed(7)
    est = (0.06 * torch.randn(N, 9, generator=g)).to(dev)
    scores = torch.rand(N, generator=g).to(dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    origin = torch.from_numpy(scene.origin)
    nv = []
    for rep in range(a.reps):
        last = rep == a.reps - 1
        if last:
            _lib.TIMERS = _lib.KernelTimers()
            if a.profile:
                torch.cuda.synchronize()
                torch.cuda.profiler.start()
        for b in frames:
            if not a.no_flush:
                flush.fill_(rep)
            depth = b['tof_depth']
            vals = ex.forward(depth, b['extrinsics'], b['intrinsics'], tsdf, wvol, origin, scene.resolution)
            filt = torch.where(b['mask'], depth, torch.zeros_like(depth)).reshape(-1)
            upd = FrameUpdate(ray=vals['ray'], filtered_depth=filt, est=est, tail=7, clamp=0.1,
                              semantics=b['semantic_gt'].reshape(-1), scores=scores)
            if not a.no_flush:
                flush.fill_(rep + 1)
            integ.forward(upd, tsdf, wvol, sc, ids)
            if last:
                nv.append(int((filt != 0).sum()))
        torch.cuda.synchronize()
        if last and a.profile:
            torch.cuda.profiler.stop()
    t = _lib.TIMERS
    e_ms, i_ms = t.mean_ms('extract'), t.mean_ms('integrate')
    peak = 6548.8
    eb, ib = 364.0 * N, 817.0 * float(np.mean(nv))
    print('frames %d  %dx%d -> %d^3  mean valid rays %.0f' % (len(frames), h, w, G, np.mean(nv)))
    print('extract   %8.1f us  %7.1f GB/s algorithmic  %.3f of HBM peak' % (1e3 * e_ms, eb / e_ms / 1e6, eb / e_ms / 1e6 / peak))
    print('integrate %8.1f us  %7.1f GB/s algorithmic  %.3f of HBM peak' % (1e3 * i_ms, ib / i_ms / 1e6, ib / i_ms / 1e6 / peak))
    per = [a_.elapsed_time(b_) * 1e3 for a_, b_ in t.events['integrate']]
    print('integrate per frame us:', ' '.join('%.0f' % x for x in per))
    print('touched voxels (weight>0): %d' % int((wvol > 0).sum()))


if __name__ == '__main__':
    main()
