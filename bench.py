#!/usr/bin/env python
"""bench.py -- fused frames/s at 240x320 RGB-D into a 256^3 grid (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one RGB-D frame through the whole per-frame path of BASELINE.json configs[1]:
AdapNet++ (stage 2, 30 classes) -> softmax/max -> Extractor -> FusionNet_v3 (semantic head on)
-> Integrator (TSDF + semantics), via the public `Pipeline.fuse(batch, database, device)`.
Data are synthetic (analytic SDF room, seeded) and the networks are random-init (seed 1911):
there is no dataset, checkpoint or network in this environment.

Own arm (default): one process per GPU, scenes sharded one-per-rank (4 scenes per rank,
rotated every frame so the voxel working set exceeds L2), no data-path collective.
  value  = frames/s with the frame tensors already resident in HBM
  e2e    = frames/s through the streaming call (stream.FrameStream.submit -> Pipeline.fuse) with HOST (pinned) frame
           tensors: H2D of image+depth+mask every step and a D2H read of a step's scalar result every step, overlapped
           with the neighbouring frames' kernels (depth-2 ring, AdapNet++ of frame i+1 next to the fusion of frame i);
           e2e.value_synchronous = the same with one blocking Pipeline.fuse + .item() per frame
  roofline = the kernels the step spends most of its own-kernel time in: the tcgen05 convolutions (csrc/ojdf_conv_ss.cu,
           ojdf_conv_tc.cu, ojdf_conv_chain.cu) of the FusionNet stack -- algorithmic conv FLOPs of FusionNet_v3(sem)
           (78.15 GFLOP per 240x320 frame) / the CUDA-event time of the engine forward,
           against the measured dense bf16 tensor peak (sustained figure: the kernel is timed inside
           a long step); roofline_adapnet = the same for AdapNet++ stage 2 (59.1 GFLOP, + ojdf_conv_wt.cu);
           roofline_integrate / roofline_extract = algorithmic bytes (817 B per valid
           ray / 364 B per ray) / CUDA-event time of those calls, against the measured HBM peak
  cpu_baseline = the unmodified reference's Pipeline.fuse on the host CPU (kind "reference"; the CPU port --
           oracle C for extract/integrate + the same torch modules on CPU -- is reported beside it as
           cpu_baseline_port) on a bounded sample, rank 0, N=1 only
  parity (--parity FRAMES) = F1 / mIoU / iou / acc of the CUDA path next to the CPU port after FRAMES frames

Reference arm (--impl reference): the UNMODIFIED reference (git-ignored baseline/_ref, installed by
baseline/install_ref.py; driven by baseline/harness.py) on the host cores, same metric/config; the CPU port only if
that tree is missing.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from online_joint_depthfusion_and_semantic_b200 import _lib  # noqa: E402
from online_joint_depthfusion_and_semantic_b200.config import fusion_config  # noqa: E402
from online_joint_depthfusion_and_semantic_b200.synthetic import SyntheticScene  # noqa: E402

H, W, GRID, N_CLASSES = 240, 320, 256, 30
# dram__bytes_read.sum + dram__bytes_write.sum per frame of each stage's kernels, from the committed `ncu --set full`
# capture of this very command (profiles/r2_*; filled in after each kernel change, None = not captured for this build)
TRAFFIC = {'fusionnet': None, 'integrate': None, 'extract': None, 'adapnet': None}
_TRAFFIC_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'r2_traffic.json')
if os.path.exists(_TRAFFIC_FILE):                      # written by profiles/make_traffic.py from the committed launch list
    try:
        _t = json.load(open(_TRAFFIC_FILE))
        TRAFFIC = {k: (_t[k]['dram_bytes_per_frame'] if k in _t else None) for k in TRAFFIC}
    except (ValueError, KeyError, TypeError):
        pass
SCENES_PER_RANK, FRAMES_PER_SCENE = 4, 6
METRIC = 'fused_frames_per_second_240x320_into_256cube'
WORKLOAD = 'configs[1]: synthetic Replica-like room, 256^3 grid, 240x320 RGB-D, AdapNet++(stage2,30cls)+FusionNet_v3(sem)+extract+integrate'
PRECISION = 'parity'


def select_config(name):
    """'headline' = BASELINE.json configs[1] (the metric's configuration, default); 'fast480' = configs[2]: 480x640 frames
    into a 512^3 grid with the tensor-core convolutions in the `fast` precision mode (1xTF32 instead of 3xTF32)."""
    global H, W, GRID, SCENES_PER_RANK, METRIC, WORKLOAD, PRECISION
    if name == 'fast480':
        H, W, GRID, SCENES_PER_RANK, PRECISION = 480, 640, 512, 2, 'fast'
        METRIC = 'fused_frames_per_second_480x640_into_512cube_fast_precision'
        WORKLOAD = ('configs[2]: synthetic Replica-like room, 512^3 grid, 480x640 RGB-D, AdapNet++(stage2,30cls)+FusionNet_v3(sem) '
                    'with 1xTF32 tensor-core convolutions (precision mode fast) + extract + integrate')


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), float(d.get('bf16_tflops_sustained', d.get('bf16_tflops', 1400.0))), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1400.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.th.join(timeout=2)
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
class SceneSet:
    """`dataset` for modules.database.Database: analytic scenes with GT grids."""

    def __init__(self, scenes, device):
        self._scenes = {s.name: s for s in scenes}
        self.scenes = list(self._scenes)
        self.device = device

    def get_grid(self, name, truncation, semantic_grid):
        from online_joint_depthfusion_and_semantic_b200.modules.database import Voxelgrid
        s = self._scenes[name]
        sdf, lab = s.gt_volumes(device=self.device, truncation=truncation)
        g = Voxelgrid(s.resolution); g.from_array(sdf, s.bbox)
        l = Voxelgrid(s.resolution); l.from_array(lab, s.bbox)
        return (g, l)


def build_world(device, rank, h=None, w=None, grid=None, scenes_per_rank=None, frames=FRAMES_PER_SCENE,
                render_device=None, strategy='predict'):
    h, w, grid = h or H, w or W, grid or GRID                # the selected configuration (select_config) unless given
    scenes_per_rank = scenes_per_rank or SCENES_PER_RANK
    from online_joint_depthfusion_and_semantic_b200.config import Config
    from online_joint_depthfusion_and_semantic_b200.modules.database import Database
    from online_joint_depthfusion_and_semantic_b200.modules.pipeline import Pipeline
    cfg = fusion_config(h, w, semantics='class30', semantic_strategy=strategy, use_semantics=True,
                        n_classes=N_CLASSES, stage=2, device=str(device))
    torch.manual_seed(1911)
    pipe = Pipeline(cfg)
    gen = torch.Generator().manual_seed(1911)
    for m in pipe.modules():                       # non-trivial BN statistics, like a trained model
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=gen))
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=gen))
    pipe = pipe.to(device).eval()
    scenes = [SyntheticScene(name='scene%d_%d' % (rank, i), grid=grid, h=h, w=w, n_frames=frames,
                             seed=rank * scenes_per_rank + i, intrinsics='pinhole')
              for i in range(scenes_per_rank)]
    db = Database(SceneSet(scenes, device), Config(device=device, implementation='efficient', init_value=0.1,
                                                   semantics='class30', semantic_grid=True, n_classes=N_CLASSES))
    rd = render_device or device
    host_frames = []
    for f in range(frames):
        for s in scenes:                            # scene rotates fastest: consecutive frames hit different volumes
            b = s.frame(f, device=rd)
            hb = {k: (v.cpu().pin_memory() if torch.is_tensor(v) and torch.cuda.is_available() else
                      (v.cpu() if torch.is_tensor(v) else v)) for k, v in b.items()}
            host_frames.append(hb)
    return cfg, pipe, db, host_frames


_DEVICE_KEYS = ('image', 'tof_depth', 'mask')


def to_device_frame(hb, device):
    """H2D of the per-frame tensors the kernels read; the pose (100 bytes) stays on the host."""
    out = dict(hb)
    for k in _DEVICE_KEYS:
        out[k] = hb[k].to(device, non_blocking=True)
    return out


def h2d_bytes(hb):
    return int(sum(hb[k].numel() * hb[k].element_size() for k in _DEVICE_KEYS))


class ResultTap:
    """Captures the step's scalar result (mean |tsdf update| of the frame) from Pipeline._fusion."""

    def __init__(self, pipe):
        self.value = None
        inner = pipe._fusion

        def tapped(inputs, values, **kw):
            est = inner(inputs, values, **kw)
            self.value = est.abs().mean()
            return est
        pipe._fusion = tapped


def run_own(args):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: the own arm needs a CUDA device (no CPU fallback); use --impl reference for the CPU port')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    torch.backends.cudnn.allow_tf32 = False          # the reference computes its convs in fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    _lib.lib()

    cfg, pipe, db, host_frames = build_world(device, rank)
    pipe.set_precision(PRECISION)
    tap = ResultTap(pipe)
    dev_frames = [to_device_frame(hb, device) for hb in host_frames]
    torch.cuda.synchronize()
    nf = len(host_frames)

    def step_resident(i):
        pipe.fuse(dict(dev_frames[i % nf]), db, device)

    def step_e2e_sync(i):
        b = to_device_frame(host_frames[i % nf], device)
        pipe.fuse(b, db, device)
        return float(tap.value.item())              # D2H read of the step's result (4 bytes) -> also a sync

    # the public streaming call (stream.py): H2D of frame i+1 and the read-back of frame i-1 overlap the kernels of frame i
    from online_joint_depthfusion_and_semantic_b200.stream import FrameStream
    fstream = FrameStream(pipe, db, device, result_fn=lambda: tap.value, depth=2, keys=_DEVICE_KEYS)
    e2e_results = []

    def step_e2e(i):
        r = fstream.submit(host_frames[i % nf])     # every step: H2D of this frame (pinned -> device), D2H of a finished frame's result
        if r is not None:
            e2e_results.append(r)

    def finish_e2e():
        e2e_results.extend(fstream.flush())         # the last frames' results are read inside the timed region too

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, with_clocks=False, finish=None):
        with torch.no_grad():
            for i in range(warmup):
                fn(i)
            if finish is not None:
                finish()
            barrier()
            sampler = ClockSampler(local) if with_clocks else None
            if sampler:
                sampler.__enter__()
            l0 = _lib.launch_count()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if args.ncu_range and with_clocks:
                torch.cuda.profiler.start()          # ncu --profile-from-start off: capture only the timed steps
            a.record()
            for i in range(steps):
                fn(warmup + i)
            if finish is not None:
                finish()
            b.record()
            barrier()
            if args.ncu_range and with_clocks:
                torch.cuda.profiler.stop()
            launches = _lib.launch_count() - l0
            if sampler:
                sampler.__exit__()
        ms = torch.tensor([a.elapsed_time(b)], device=device, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), launches, (sampler.summary() if sampler else None)

    # --- value: frames resident in HBM, kernel timers on for the roofline leg
    _lib.TIMERS = _lib.KernelTimers()
    ms_total, launches, clocks = timed(step_resident, args.steps, args.warmup, with_clocks=True)
    timers, _lib.TIMERS = _lib.TIMERS, None
    torch.cuda.synchronize()
    n_timed = args.steps
    def stage(name):
        ev = timers.events.get(name, [])[-n_timed:]
        return float(np.mean([a.elapsed_time(b) for a, b in ev])) if ev else None
    ext_ms, int_ms, plan_ms, rays_ms = stage('extract'), stage('integrate'), stage('integrate_plan') or 0.0, stage('rays') or 0.0
    stages = {k: stage(k) for k in ('adapnet', 'rays', 'extract', 'fusionnet', 'integrate_plan', 'integrate')}
    # --- e2e: host frames, H2D + D2H inside the timed region
    ms_e2e = ms_e2e_sync = float('nan')
    if not args.skip_e2e:
        ms_e2e, _, _ = timed(step_e2e, args.steps, max(3, args.warmup // 2), finish=finish_e2e)
        assert len(e2e_results) >= args.steps and all(np.isfinite(e2e_results))
        ms_e2e_sync, _, _ = timed(step_e2e_sync, max(20, args.steps // 4), 3)
        ms_e2e_sync *= args.steps / max(20, args.steps // 4)

    fps = world * args.steps / (ms_total / 1e3)
    fps_e2e = world * args.steps / (ms_e2e / 1e3)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    n_rays = H * W
    nv = float(np.mean([int((hb['mask'] & (hb['tof_depth'] != 0)).sum()) for hb in host_frames]))
    peak, tpeak, peak_src = peaks()
    int_bytes, ext_bytes = 817.0 * nv, 364.0 * n_rays
    int_total = int_ms + plan_ms
    roof_int = {'kernel': 'ojdf_integrate_plan (count + offsets + scatter; side stream, overlaps the networks) + '
                          'ojdf_integrate_apply (apply_short + apply_long; the only part after FusionNet)', 'bound': 'hbm',
                'achieved': int_bytes / (int_total * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                'frac': int_bytes / (int_total * 1e-3) / 1e9 / peak, 'traffic': TRAFFIC.get('integrate') if PRECISION == 'parity' else None,
                'peak_source': peak_src,
                'algorithmic_bytes_per_launch': int_bytes, 'ms_per_launch': float(int_total),
                'ms_plan_side_stream': float(plan_ms), 'ms_apply_critical_path': float(int_ms)}
    ext_total = ext_ms + rays_ms
    roof_ext = {'kernel': 'ojdf_rays + ojdf_gather (extract_kernel: per-ray records, then the fused gather that also writes '
                          'FusionNet\'s input rows)', 'bound': 'hbm',
                'achieved': ext_bytes / (ext_total * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                'frac': ext_bytes / (ext_total * 1e-3) / 1e9 / peak, 'traffic': TRAFFIC.get('extract') if PRECISION == 'parity' else None, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': ext_bytes, 'ms_per_launch': float(ext_total)}
    fn_flop = 78.15e9 * (H * W) / 76800.0            # FusionNet_v3 (semantic head on), convolutions only, 2*MAC
    fn_ms = stages['fusionnet']
    roof_conv = {'kernel': 'the tcgen05 kind::tf32 convolution kernels (3xTF32 split precision) over the FusionNet_v3 stack: ss::conv_ss_kernel '
                           '(dense blocks), tc::conv_tc_kernel (vortex 1x1 / dilated 3x3), chain::conv_chain_kernel (vortex tails, Pred stack): '
                           '21 launches per frame + 12 small pooling / bias launches inside the same bracket',
                 'bound': 'tensor', 'achieved': fn_flop / (fn_ms * 1e-3) / 1e12, 'peak': tpeak, 'unit': 'TFLOP/s',
                 'frac': fn_flop / (fn_ms * 1e-3) / 1e12 / tpeak, 'traffic': TRAFFIC.get('fusionnet') if PRECISION == 'parity' else None,
                 'peak_source': peak_src + ', dense bf16 sustained; the kernel issues 3 tf32 MMAs per algorithmic MAC '
                                           '(tf32 dense peak is half of bf16), so 1/6 of this peak is its arithmetic ceiling',
                 'algorithmic_flop_per_frame': fn_flop, 'ms_per_frame': float(fn_ms)} if fn_ms else None
    an_flop = 59.1e9 * (H * W) / 76800.0             # AdapNet++ stage 2 (both encoders, eASPP, SSMA, decoder), convolutions only, 2*MAC
    an_ms = stages['adapnet']
    roof_adap = {'kernel': 'the same kernels + wt::conv_wt_kernel (small maps: output channels as M, the image as N) over AdapNet++ stage 2, '
                           'replayed as one CUDA graph (stem, softmax / arg-max and split-K reductions inside the same bracket)',
                 'bound': 'tensor', 'achieved': an_flop / (an_ms * 1e-3) / 1e12, 'peak': tpeak, 'unit': 'TFLOP/s',
                 'frac': an_flop / (an_ms * 1e-3) / 1e12 / tpeak, 'traffic': TRAFFIC.get('adapnet') if PRECISION == 'parity' else None,
                 'peak_source': peak_src + ', dense bf16 sustained (1/6 of it is the 3xTF32 arithmetic ceiling)',
                 'algorithmic_flop_per_frame': an_flop, 'ms_per_frame': float(an_ms)} if an_ms else None
    line = {
        'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64 ray geometry / f32 accumulation / f16+u8 volumes; FusionNet + AdapNet++ convolutions: '
                 + ('fp32 via 3xTF32 tcgen05 (own kernels, ~1e-6 of fp32)' if PRECISION == 'parity' else '1xTF32 tcgen05 (own kernels, precision mode fast, ~1e-3)')
                 + '; AdapNet++ 7x7 stem: fp32 FMA (own kernel)',
        'data': 'synthetic (analytic SDF room, seeded; random-init networks seed 1911)',
        'config': {'workload': WORKLOAD, 'frame': [H, W], 'grid': GRID, 'scenes_per_gpu': SCENES_PER_RANK,
                   'sharding': 'scenes one-per-rank, no collective',
                   'l2': 'inputs larger than L2: %d scenes x 117 MB of volumes rotated every frame + >1 GB of network activations per frame' % SCENES_PER_RANK},
        'e2e': {'value': fps_e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d_bytes(host_frames[0]), 'd2h_bytes_per_step': 4,
                'call': 'stream.FrameStream.submit(host_batch) -> Pipeline.fuse: pinned host frames, depth-2 ring (H2D of frame i+1 and '
                        'the D2H read of frame i-1 overlap the kernels of frame i; every copy and read is inside the timed region)',
                'value_synchronous': world * args.steps / (ms_e2e_sync / 1e3),
                'synchronous_call': 'Pipeline.fuse(host batch) + .item() of the result every step (no overlap between frames)'},
        'gpu_launches': int(launches), 'clocks': clocks,
        'roofline': roof_conv, 'roofline_adapnet': roof_adap, 'roofline_integrate': roof_int, 'roofline_extract': roof_ext,
        'stage_ms': stages,
    }
    if world == 1 and not args.no_cpu_baseline:
        port = cpu_port_fps(steps=3, warmup=1)
        ref = reference_fps(steps=3, warmup=1)
        line['cpu_baseline'] = ref if ref is not None else port
        line['cpu_baseline_port'] = port
    if world == 1 and args.parity:
        sys.path.insert(0, os.path.join(ROOT, 'tools'))
        import parity_report
        torch.cuda.empty_cache()
        par = parity_report.parity(frames=args.parity, h=H, w=W, grid=GRID)
        line['parity'] = {'frames': par['frames'], 'protocol': par['protocol'], 'cuda': par['cuda'], 'cpu_port': par['cpu_port'],
                          'max_abs_diff_points': par['max_abs_diff_points'], 'volumes': par['volumes'],
                          'within_half_point': bool(par['max_abs_diff_points'] <= 0.5)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def cpu_port_fps(steps, warmup, verbose=False):
    """The reference's CPU path restated: oracle C (extract / integrate, all host threads it can use)
    + the same torch modules on CPU for AdapNet++ and FusionNet_v3.  One step = one frame."""
    from oracle import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oracle.set_threads(cores)
    dev = torch.device('cpu')
    n_scenes = 1
    cfg, pipe, db, frames = build_world(dev, 0, scenes_per_rank=n_scenes, frames=max(2, min(4, steps + warmup)),
                                        render_device='cuda' if torch.cuda.is_available() else 'cpu')
    P, T = 9, 7
    vol = {k: None for k in db.scenes}
    for s in db.scenes:
        vol[s] = [db.scenes_est[s].volume.numpy().view(np.uint16), db.fusion_weights[s].numpy().view(np.uint16),
                  db.ids_est[s].volume.numpy(), db.scores[s].volume.numpy().view(np.uint16)]

    def frame(i):
        b = frames[i % len(frames)]
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
            values = {'fusion_values': torch.from_numpy(o['fusion_values'])[None], 'fusion_weights': torch.from_numpy(o['fusion_weights'])[None]}
            est = pipe._fusion(pipe._prepare_fusion_input(depth, values, sem), values)
            oracle.integrate_frame(world, filt.reshape(-1).numpy(), est[0].contiguous().numpy(), E[:3, 3], db.origin[scene].numpy(), res,
                                   tsdf, wvol, tail=T, clampv=0.1, pix_ids=sem.reshape(-1).to(torch.uint8).numpy(),
                                   pix_scores=scores.reshape(-1).numpy(), ids_vol=ids, scores_vol=sc, do_sem=True)

    for i in range(warmup):
        frame(i)
    t0 = time.perf_counter()
    for i in range(steps):
        frame(warmup + i)
    dt = time.perf_counter() - t0
    return {'value': steps / dt, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
            'sample': '%d frames (after %d warm-up) of the same workload on one scene: oracle C extract+integrate '
                      '(pthreads) + torch-CPU AdapNet++/FusionNet_v3, %d threads' % (steps, warmup, cores),
            's_per_frame': dt / steps}


def reference_fps(steps, warmup):
    """The UNMODIFIED reference's own Pipeline.fuse (baseline/_ref, or /root/reference in the build container) on the host
    CPU with every thread torch can use, in a child process (its `modules` / `utils` packages must not leak into this
    one).  None if the reference tree is not there."""
    from baseline import harness
    if harness.ref_root() is None:
        return None
    cores = os.cpu_count() or 1
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'baseline', 'harness.py'), 'fps', '--frames', str(steps), '--h', str(H),
                        '--w', str(W), '--grid', str(GRID), '--threads', str(cores)], capture_output=True, text=True, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith('HARNESS_RESULT ')]
    if r.returncode != 0 or not lines:
        sys.stderr.write('bench.py: reference run failed, falling back to the CPU port\n' + r.stderr[-2000:] + '\n')
        return None
    d = json.loads(lines[-1][len('HARNESS_RESULT '):])
    return {'value': d['fps'], 'unit': 'frames/s', 'cores': int(d['threads']), 'kind': 'reference',
            'sample': '%d frames (after 1 warm-up) of the same workload on one scene through the unmodified reference '
                      'Pipeline.fuse (modules/pipeline.py:173-248: AdapNet++ stage 2 + Extractor + FusionNet_v3 + Integrator) '
                      'on torch-CPU, %d threads' % (d['frames'], d['threads']),
            's_per_frame': d['s_per_frame']}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = min(args.steps, 8)                      # ~3.7 s per frame for the reference, ~0.7 s for the port
    warmup = 1
    r = reference_fps(steps, warmup)
    if r is None:
        steps = min(args.steps, 20)
        r = cpu_port_fps(steps, min(args.warmup, 2))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': 1e3 * r['s_per_frame'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64/f32/f16 as the reference (CPU)', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'frame': [H, W], 'grid': GRID,
                   'note': 'kind "reference": the unmodified reference from git-ignored baseline/_ref (python baseline/install_ref.py) '
                           'driven by baseline/harness.py; kind "port" (only if that tree is missing): its CPU path restated '
                           '(oracle C + identical torch CPU modules), pinned bit-exact to reference-generated fixtures'},
        'cpu_baseline': {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': r['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def run_train(args):
    """--mode train: config 5 of BASELINE.json on synthetic frames -- every rank streams its own scenes through
    Pipeline.fuse_training (frozen AdapNet++ on the own kernels, Extractor x2, FusionNet_v3 with autograd, FusionLoss,
    Integrator test=False), gradients accumulate for 8 frames with the reference's per-iteration clip
    (train_fusion.py:166-189), then ONE NCCL all-reduce of the flat FusionNet gradient bucket and an RMSprop step.
    A step = one frame; value = frames/s over all ranks; the all-reduce is timed with CUDA events."""
    from online_joint_depthfusion_and_semantic_b200.training import FusionLoss, PolynomialLR, ShardedFusionTrainer
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --mode train needs CUDA devices')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = os.environ.get('OJDF_CUDNN_BENCHMARK', '1') != '0'   # fp32 algorithms picked by measurement
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    cfg, pipe, db, host_frames = build_world(device, rank, scenes_per_rank=2, frames=8)
    pipe.train()
    pipe._semantic_2d_network.eval()                         # frozen, as train_fusion.py:90-94
    for p_ in pipe._semantic_2d_network.parameters():
        p_.requires_grad_(False)
    opt = torch.optim.RMSprop(pipe._fusion_network.parameters(), lr=1e-5, momentum=0.9, weight_decay=0.01, eps=1e-9)
    trainer = ShardedFusionTrainer(pipe, opt, PolynomialLR(opt, max_iter=50000), FusionLoss(), accumulation_steps=8, clipping=True)
    trainer.time_collective = True
    trainer.broadcast_parameters()
    dev_frames = [to_device_frame(hb, device) for hb in host_frames]
    nf = len(dev_frames)

    def step(i):
        b = dict(dev_frames[i % nf])
        b['tof_depth'] = b['tof_depth'].clone()              # fuse_training squeezes its input in place
        loss, _ = trainer.train_frame(b, db, device)
        return loss

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    steps = max(8, args.steps // 8 * 8) if args.steps < 200 else 64          # whole accumulation windows
    for i in range(max(8, args.warmup // 8 * 8)):
        step(i)
    barrier()
    trainer.collective_events.clear()
    with ClockSampler(local) as sampler:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = None
        for i in range(steps):
            loss = step(i)
        b.record()
        barrier()
    ms = torch.tensor([a.elapsed_time(b)], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ar = [x.elapsed_time(y) * 1e3 for x, y in trainer.collective_events]
    # the collective alone: the same 2.29 MB bucket all-reduced back to back after a barrier (the in-loop figure above
    # also contains the wait for the slowest rank to reach the collective)
    iso = None
    if dist is not None:
        bucket = trainer._flat_bucket().clone()
        for _ in range(5):
            dist.all_reduce(bucket)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            dist.all_reduce(bucket)
        e1.record()
        torch.cuda.synchronize()
        iso_t = torch.tensor([e0.elapsed_time(e1) * 1e3 / 50], device=device, dtype=torch.float64)
        dist.all_reduce(iso_t, op=dist.ReduceOp.MAX)
        iso = float(iso_t.item())
    nparam = sum(p_.numel() for p_ in trainer.params)
    if rank == 0:
        print(json.dumps({
            'mode': 'train', 'metric': 'online_training_frames_per_second_240x320_into_256cube', 'value': world * steps / (float(ms.item()) / 1e3),
            'unit': 'frames/s', 'n_gpus': world, 'steps': steps, 'warmup': 8, 'ms_per_step': float(ms.item()) / steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'fp32 (FusionNet forward/backward: PyTorch autograd, cuDNN, TF32 off); frozen AdapNet++ + extract + integrate: own kernels',
            'data': 'synthetic', 'config': {'workload': 'configs[4]: fuse_training on scene-sharded streams, 8-frame accumulation, '
                                                       'one NCCL all-reduce of the flat FusionNet gradient per optimiser step',
                                            'frame': [H, W], 'grid': GRID, 'accumulation_steps': 8},
            'allreduce': {'bytes': 4 * nparam, 'calls': len(ar), 'us_in_loop_mean_incl_rank_skew': float(np.mean(ar)) if ar else None,
                          'us_in_loop_min': float(np.min(ar)) if ar else None, 'us_isolated_back_to_back': iso,
                          'bus_GBps_isolated': (4 * nparam * 2 * (world - 1) / world / (iso * 1e-6) / 1e9) if iso else None},
            'loss': float(loss), 'clocks': sampler.summary()}))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--mode', default='fuse', choices=['fuse', 'train'],
                    help="fuse (default): the headline inference metric; train: BASELINE.json configs[4], online training with the NCCL gradient all-reduce")
    ap.add_argument('--config', default='headline', choices=['headline', 'fast480'],
                    help='headline: BASELINE.json configs[1] (default); fast480: configs[2], 480x640 -> 512^3, precision mode fast')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--parity', type=int, default=0, metavar='FRAMES',
                    help='also fuse FRAMES frames through the CUDA path and the CPU port and report the metric parity (N=1)')
    ap.add_argument('--ncu-range', action='store_true', help='bracket the timed value-steps with cudaProfilerStart/Stop')
    ap.add_argument('--skip-e2e', action='store_true', help='(profiling runs only) skip the e2e leg')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    select_config(args.config)
    if args.impl == 'reference':
        run_reference(args)
    elif args.mode == 'train':
        run_train(args)
    else:
        run_own(args)


if __name__ == '__main__':
    main()
