#!/usr/bin/env python
"""F-score / mIoU parity report (BASELINE.json `metric`): fuse the same synthetic frames
  (a) through the CUDA pipeline (`Pipeline.fuse`: own extract / integrate / conv kernels), and
  (b) through the CPU port of the reference path (oracle C extract + integrate, the same torch modules on CPU),
then evaluate both sets of volumes with `Database.evaluate` / `evaluate_semantics` (after `filter`,
`filter_semantics(5)`, exactly as test_fusion.py:82-108) and print the metrics side by side.

    python tools/parity_report.py [--frames 8] [--h 120 --w 160 --grid 64] [--cpu-only]

Test infrastructure: this is one of the places that may call oracle/ (it is the checker here)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def fuse_cpu_port(pipe, db, frames):
    """The frame loop of bench.cpu_port_fps: oracle C for the two memory-bound ops, torch-CPU networks."""
    from oracle import oracle
    dev = torch.device('cpu')
    vol = {s: [db.scenes_est[s].volume.numpy().view(np.uint16), db.fusion_weights[s].numpy().view(np.uint16),
               db.ids_est[s].volume.numpy(), db.scores[s].volume.numpy().view(np.uint16)] for s in db.scenes}
    for b in frames:
        pipe.device = dev
        pipe._shape = b['image'].shape
        scene = b['frame_id'][0].split('/')[0]
        tsdf, wvol, ids, sc = vol[scene]
        with torch.no_grad():
            scores, sem = pipe._semantic_frame(b, as_uint8=False)
            depth = b['tof_depth']
            filt = torch.where(b['mask'], depth, torch.zeros_like(depth))
            E = b['extrinsics'][0].numpy()
            Kinv = b['intrinsics'][0].float().inverse().numpy()
            world = oracle.unproject(depth[0].numpy(), Kinv, E)
            res = db.resolution[scene]
            o = oracle.extract(world, E[:3, 3], db.origin[scene].numpy(), res, tsdf, wvol)
            values = {'fusion_values': torch.from_numpy(o['fusion_values'])[None],
                      'fusion_weights': torch.from_numpy(o['fusion_weights'])[None]}
            est = pipe._fusion(pipe._prepare_fusion_input(depth, values, sem), values)
            oracle.integrate_frame(world, filt.reshape(-1).numpy(), est[0].contiguous().numpy(), E[:3, 3],
                                   db.origin[scene].numpy(), res, tsdf, wvol, tail=7, clampv=0.1,
                                   pix_ids=sem.reshape(-1).to(torch.uint8).numpy(), pix_scores=scores.reshape(-1).numpy(),
                                   ids_vol=ids, scores_vol=sc, do_sem=True)
        db.state[scene] = True


def report(db):
    db.filter()
    db.filter_semantics(value=5)
    geo, _ = db.evaluate(mode='test')
    sem, _ = db.evaluate_semantics(mode='test')
    return {**geo, **sem}


def parity(frames=100, h=240, w=320, grid=256, cuda=True, strategy='predict'):
    """Fuse `frames` synthetic frames of one scene through both paths (bottleneck dropout off on both sides, as
    SURVEY.md 0.6 requires for a deterministic comparison) and return the test_fusion.py:82-108 metrics side by side."""
    assert h % 16 == 0 and w % 16 == 0, 'AdapNet++ needs h, w = 0 mod 16'
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import oracle
    oracle.set_threads(os.cpu_count() or 1)
    out = {'frames': frames, 'frame': [h, w], 'grid': grid, 'semantic_strategy': strategy,
           'protocol': 'filter(2.0) + filter_semantics(5) + evaluate / evaluate_semantics, as test_fusion.py:82-108'}
    cuda = cuda and torch.cuda.is_available()
    _, pipe_c, db_c, host_frames = bench.build_world(torch.device('cpu'), 0, h=h, w=w, grid=grid, scenes_per_rank=1,
                                                     frames=frames, render_device='cuda' if cuda else 'cpu', strategy=strategy)
    if pipe_c._semantic_2d_network is not None:
        pipe_c._semantic_2d_network.set_bottleneck_dropout(False)       # deterministic on both sides
    fuse_cpu_port(pipe_c, db_c, host_frames)
    if not cuda:
        out['cpu_port'] = report(db_c)
    if cuda:
        dev = torch.device('cuda', 0)
        _, pipe_g, db_g, _ = bench.build_world(dev, 0, h=h, w=w, grid=grid, scenes_per_rank=1, frames=1, strategy=strategy)
        if pipe_g._semantic_2d_network is not None:
            pipe_g._semantic_2d_network.set_bottleneck_dropout(False)
        with torch.no_grad():
            for hb in host_frames:
                pipe_g.fuse(bench.to_device_frame(hb, dev), db_g, dev)
        torch.cuda.synchronize()
        # raw volume agreement BEFORE any filtering (report() filters in place).  The two paths differ by ~1e-6 in the network outputs (3xTF32 vs
        # fp32 summation order); over `frames` frames the fp16 running means round differently now and then, so many
        # voxels end up a few fp16 ulps apart -- what matters is how far, and that the metrics do not move.
        s = db_g.scenes[0]
        sc = db_c.scenes[0]
        t_g, t_c = db_g.scenes_est[s].volume.cpu().float(), db_c.scenes_est[sc].volume.float()
        i_g, i_c = db_g.ids_est[s].volume.cpu(), db_c.ids_est[sc].volume
        touched_m = db_c.fusion_weights[sc] > 0
        touched = int(touched_m.sum())
        d = (t_g - t_c).abs()[touched_m]
        ulp = 2.0 ** -14                                                 # fp16 spacing just below 0.125 (|tsdf| <= 0.1)
        out['volumes'] = {'touched_voxels': touched,
                          'tsdf_voxels_differing': int((d > 0).sum()),
                          'tsdf_voxels_differing_by_more_than_2_fp16_ulp': int((d > 2 * ulp).sum()),
                          'tsdf_max_abs_diff': float(d.max()) if touched else 0.0,
                          'tsdf_mean_abs_diff': float(d.mean()) if touched else 0.0,
                          'label_voxels_differing': int((i_g != i_c)[touched_m].sum()),
                          'label_agreement': float((i_g == i_c)[touched_m].float().mean()) if touched else 1.0}
        out['cpu_port'] = report(db_c)
        out['cuda'] = report(db_g)
        out['abs_diff'] = {k: abs(out['cuda'][k] - out['cpu_port'][k]) for k in out['cuda']}
        out['max_abs_diff_points'] = 100.0 * max(out['abs_diff'][k] for k in ('iou', 'acc', 'f1', 'Mean IoU', 'Mean Acc')
                                                 if k in out['abs_diff'])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=8)
    ap.add_argument('--h', type=int, default=120)
    ap.add_argument('--w', type=int, default=160)
    ap.add_argument('--grid', type=int, default=64)
    ap.add_argument('--cpu-only', action='store_true')
    ap.add_argument('--out', default='')
    a = ap.parse_args()
    out = parity(a.frames, a.h, a.w, a.grid, cuda=not a.cpu_only)
    txt = json.dumps(out, indent=1)
    if a.out:
        with open(a.out, 'w') as f:
            f.write(txt + '\n')
    print(txt)


if __name__ == '__main__':
    main()
