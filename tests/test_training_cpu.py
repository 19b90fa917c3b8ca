"""CPU tests of the training-side host logic: the restated FusionLoss against the formula the
reference implements, and the scene-sharded gradient all-reduce over gloo with world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from online_joint_depthfusion_and_semantic_b200.training import FusionLoss, PolynomialLR, ShardedFusionTrainer


def test_fusion_loss_matches_torch14_formula():
    torch.manual_seed(0)
    est = torch.randn(1, 37, 9, requires_grad=True)
    tgt = torch.randn(1, 37, 9)
    loss = FusionLoss(w_l1=1., w_l2=10., w_cos=0.1)(est, tgt)
    x1, x2 = torch.sign(est).reshape(1, 9, 37), torch.sign(tgt).reshape(1, 9, 37)
    # torch >= 1.5 still implements the same maths for 2-D inputs: feed it column by column
    cos = torch.stack([torch.nn.functional.cosine_embedding_loss(x1[0, :, j][None], x2[0, :, j][None], torch.ones(1))
                       for j in range(37)]).mean()
    ref = (est - tgt).abs().mean() + 10 * ((est - tgt) ** 2).mean() + 0.1 * cos
    assert torch.allclose(loss, ref, rtol=1e-6, atol=1e-7)
    loss.backward()
    assert torch.isfinite(est.grad).all()
    # no valid rays -> constant 1 without a graph (utils/loss.py:81-82)
    empty = FusionLoss()(torch.zeros(1, 0, 9), torch.zeros(1, 0, 9))
    assert float(empty) == 1.0 and empty.grad_fn is None


def test_polynomial_lr():
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    sch = PolynomialLR(opt, max_iter=100)
    for _ in range(10):
        opt.step(); sch.step()
    assert abs(opt.param_groups[0]['lr'] - (1 - 10 / 100.0) ** 0.9) < 1e-12


class _TinyFusion(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(9, 9)
        self.b = torch.nn.Linear(9, 9)


class _FakePipeline:
    """fuse_training stand-in: a differentiable map of per-rank data through the shared net."""

    def __init__(self):
        torch.manual_seed(1234)
        self._fusion_network = _TinyFusion()

    def fuse_training(self, batch, database, device):
        x = batch['x']
        y = self._fusion_network.b(torch.tanh(self._fusion_network.a(x)))
        return {'tsdf_est': y, 'tsdf_fused': y, 'tsdf_target': batch['t']}


def _rank_data(rank, frames):
    g = torch.Generator().manual_seed(100 + rank)
    return [{'x': torch.randn(1, 11, 9, generator=g), 't': 0.05 * torch.randn(1, 11, 9, generator=g)} for _ in range(frames)]


def _worker(rank, world, port, frames, acc, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        pipe = _FakePipeline()
        opt = torch.optim.RMSprop(pipe._fusion_network.parameters(), lr=1e-2, momentum=0.9, eps=1e-9)
        tr = ShardedFusionTrainer(pipe, opt, accumulation_steps=acc, clipping=True)
        tr.broadcast_parameters()
        steps = 0
        for i, b in enumerate(_rank_data(rank, frames)):
            _, stepped = tr.train_frame(b, None, 'cpu', last=(i == frames - 1))
            steps += stepped
            if stepped and i == 1:
                # what the reference loop does around evaluation (train_fusion.py:194): set_to_none=True detaches the
                # gradient views from the bucket; the trainer must re-attach them or the ranks drift apart silently
                opt.zero_grad()
                assert all(p.grad is None for p in pipe._fusion_network.parameters())
        flat = torch.cat([p.detach().reshape(-1) for p in pipe._fusion_network.parameters()])
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        if rank == 0:
            out.put((steps, [g.numpy().copy() for g in gathered]))     # by value: the worker may exit first
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _expected(world, frames, acc):
    """Same schedule in one process: per-rank accumulated+clipped grads, averaged, one RMSprop step."""
    pipe = _FakePipeline()
    net = pipe._fusion_network
    opt = torch.optim.RMSprop(net.parameters(), lr=1e-2, momentum=0.9, eps=1e-9)
    crit = FusionLoss()
    params = list(net.parameters())
    data = [_rank_data(r, frames) for r in range(world)]
    acc_g = [[torch.zeros_like(p) for p in params] for _ in range(world)]
    for i in range(frames):
        for r in range(world):
            for j, p in enumerate(params):
                p.grad = acc_g[r][j].clone()
            out = pipe.fuse_training(data[r][i], None, 'cpu')
            crit(out['tsdf_fused'], out['tsdf_target']).backward()
            torch.nn.utils.clip_grad_norm_(params, max_norm=1., norm_type=2)
            acc_g[r] = [p.grad.clone() for p in params]
        if (i + 1) % acc == 0 or i == frames - 1:
            for j, p in enumerate(params):
                p.grad = sum(acc_g[r][j] for r in range(world)) / world
            opt.step()
            acc_g = [[torch.zeros_like(p) for p in params] for _ in range(world)]
    return torch.cat([p.detach().reshape(-1) for p in params])


@pytest.mark.timeout(120)
def test_sharded_trainer_gloo_world2():
    world, frames, acc = 2, 5, 2
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, frames, acc, out)) for r in range(world)]
    for p in procs:
        p.start()
    steps, gathered = out.get(timeout=100)
    gathered = [torch.from_numpy(g) for g in gathered]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert steps == 3                                           # frames 2, 4 and the final partial window
    assert torch.equal(gathered[0], gathered[1])                # replicas stay identical
    exp = _expected(world, frames, acc)
    assert torch.allclose(gathered[0], exp, rtol=1e-5, atol=1e-7)
