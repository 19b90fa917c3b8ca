"""Row g2 / (e): the scene-sharded trainer on real hardware -- two ranks, NCCL, the real Pipeline.fuse_training.
After one accumulation window the all-reduced, averaged gradient bucket must equal the average of the two ranks'
locally accumulated (and clipped) gradients computed in a single process, and the replicas must stay bit-identical.
Needs 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import socket
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _world(rank, device):
    import bench
    cfg, pipe, db, frames = bench.build_world(device, rank, h=48, w=64, grid=48, scenes_per_rank=1, frames=4)
    pipe.train()
    pipe._semantic_2d_network.eval()
    pipe._semantic_2d_network.set_bottleneck_dropout(False)
    for m in pipe._fusion_network.modules():                # deterministic: no Dropout2d, frozen BN statistics
        if isinstance(m, torch.nn.Dropout2d):
            m.p = 0.0
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
    return pipe, db, [bench.to_device_frame(f, device) for f in frames]


def _accumulate(trainer, pipe, db, frames, device, steps):
    for i in range(steps):
        b = dict(frames[i % len(frames)])
        b['tof_depth'] = b['tof_depth'].clone()
        trainer.train_frame(b, db, device)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from online_joint_depthfusion_and_semantic_b200.training import ShardedFusionTrainer
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        pipe, db, frames = _world(rank, dev)
        opt = torch.optim.SGD(pipe._fusion_network.parameters(), lr=0.0)       # lr 0: the bucket keeps the averaged gradient
        tr = ShardedFusionTrainer(pipe, opt, accumulation_steps=4, clipping=True)
        tr.broadcast_parameters()
        # local accumulation only (3 frames), snapshot, then the 4th frame triggers the all-reduce
        seen = {}
        orig_step = opt.step

        def step(*a, **k):
            seen['avg'] = tr._flat_bucket().detach().clone()
            return orig_step(*a, **k)
        opt.step = step
        orig_ar = dist.all_reduce

        def spy(t, *a, **k):
            if t.numel() == tr._flat_bucket().numel():
                seen['local'] = t.detach().clone()
            return orig_ar(t, *a, **k)
        dist.all_reduce = spy
        _accumulate(tr, pipe, db, frames, dev, 4)
        dist.all_reduce = orig_ar
        torch.cuda.synchronize()
        gathered = [torch.zeros_like(seen['local']) for _ in range(world)]
        dist.all_gather(gathered, seen['local'])
        expect = sum(g.double() for g in gathered) / world
        err = float((seen['avg'].double() - expect).abs().max() / expect.abs().max().clamp(min=1e-30))
        flat = torch.cat([p.detach().reshape(-1) for p in pipe._fusion_network.parameters()])
        allp = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(allp, flat)
        if rank == 0:
            out.put((err, bool(torch.equal(allp[0], allp[1])), float(expect.abs().max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_nccl_gradient_allreduce_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    err, same, scale = out.get(timeout=500)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert scale > 0 and err <= 1e-6, (err, scale)            # NCCL sum / world == mean of the ranks' gradients
    assert same                                               # replicas identical after the step
