"""Online-training side of the fusion path: the loss (a18) and the scene-sharded data-parallel
optimiser step.

FusionLoss restates utils/loss.py:65-103 with the torch-1.4 `cosine_embedding_loss` semantics the
reference was written against (the stock module raises on torch >= 1.5 because the reference feeds
it 3-D tensors):
    loss = w_l1*mean|e-t| + w_l2*mean (e-t)^2 + w_cos*mean(1 - cos(sign e, sign t))
where the cosine runs along dim 1 of a RESHAPE (not a transpose) of (1,Nv,9) to (1,9,Nv)
(utils/loss.py:87-89) and cos = sum(x1*x2) / sqrt((sum x1^2 + 1e-12) * (sum x2^2 + 1e-12)).

ShardedFusionTrainer is what replaces the reference's single-GPU loop body
(train_fusion.py:166-189) when scenes are sharded one-per-GPU: every rank runs
`Pipeline.fuse_training` on its own scene stream (volumes, frames and AdapNet stay local, no
traffic), gradients accumulate locally for `accumulation_steps` frames with the reference's
every-iteration clip of the running gradient (train_fusion.py:182-183), then ONE all-reduce of the
flat FusionNet gradient bucket (2.29 MB fp32 with the semantic head) averages the ranks, followed by
the identical optimizer / scheduler step on every rank.  With world_size == 1 it is the reference's
loop body.
"""
import torch
import torch.distributed as dist
from torch import nn


class FusionLoss(nn.Module):
    def __init__(self, reduction='none', w_l1=1., w_l2=10., w_cos=0.1):
        super().__init__()
        self.lambda1 = w_l1 if w_l1 is not None else 0.
        self.lambda2 = w_l2 if w_l2 is not None else 0.
        self.lambda3 = w_cos if w_cos is not None else 0.

    def forward(self, est, target):
        if est.shape[1] == 0:                                   # no valid ray (utils/loss.py:81-82)
            return torch.ones_like(est).sum().clamp(min=1)
        b, n, p = est.shape
        x1 = torch.sign(est).reshape(b, p, n)
        x2 = torch.sign(target).reshape(b, p, n)
        eps = 1e-12
        cos = (x1 * x2).sum(1) / torch.sqrt(((x1 * x1).sum(1) + eps) * ((x2 * x2).sum(1) + eps))
        l3 = (1.0 - cos).mean()                                  # label == 1 everywhere, margin unused
        diff = est - target
        l1 = diff.abs().mean()
        l2 = (diff * diff).mean()
        return self.lambda1 * l1 + self.lambda2 * l2 + self.lambda3 * l3


class PolynomialLR(torch.optim.lr_scheduler._LRScheduler):
    """utils/schedulers.py:12-28 as it actually behaves: factor (1 - step/max_iter)^gamma every step."""

    def __init__(self, optimizer, max_iter, decay_iter=1, gamma=0.9, last_epoch=-1):
        self.max_iter, self.decay_iter, self.gamma = max_iter, decay_iter, gamma
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        factor = (1 - self.last_epoch / float(self.max_iter)) ** self.gamma
        return [base_lr * factor for base_lr in self.base_lrs]


class ShardedFusionTrainer:
    """One optimiser step per `accumulation_steps` frames; gradients all-reduced once per step."""

    def __init__(self, pipeline, optimizer, scheduler=None, criterion=None, accumulation_steps=8, clipping=True,
                 process_group=None):
        self.pipeline, self.optimizer, self.scheduler = pipeline, optimizer, scheduler
        self.criterion = criterion or FusionLoss()
        self.accumulation_steps, self.clipping, self.group = int(accumulation_steps), clipping, process_group
        self.params = [p for p in pipeline._fusion_network.parameters() if p.requires_grad]
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self._bucket = None
        self.iteration = 0
        self.time_collective = False                             # bench: CUDA events around every all-reduce
        self.collective_events = []

    def _flat_bucket(self):
        """All FusionNet gradients as views into one contiguous buffer -> a single collective.

        The views are re-checked on every call: `optimizer.zero_grad()` / `model.zero_grad()` default to
        set_to_none=True (the reference calls them around evaluation, train_fusion.py:176,194), after which autograd
        allocates fresh, unrelated .grad tensors -- all-reducing the stale bucket would then silently desynchronise the
        replicas.  A parameter whose .grad is not the expected view is re-attached (its gradient, if any, is copied in)."""
        if self._bucket is None:
            n = sum(p.numel() for p in self.params)
            self._bucket = torch.zeros(n, dtype=self.params[0].dtype, device=self.params[0].device)
            self._offsets, off = [], 0
            for p in self.params:
                self._offsets.append(off)
                off += p.numel()
        base, esz = self._bucket.data_ptr(), self._bucket.element_size()
        for p, off in zip(self.params, self._offsets):
            g = p.grad
            if g is not None and g.data_ptr() == base + off * esz and g.is_contiguous() and g.numel() == p.numel():
                continue
            view = self._bucket[off:off + p.numel()].view_as(p)
            if g is not None:
                view.copy_(g)
            else:
                view.zero_()
            p.grad = view
        return self._bucket

    def zero_grad(self):
        """Use this instead of optimizer.zero_grad(): zeroes the bucket and keeps the gradient views attached."""
        self._flat_bucket().zero_()

    def broadcast_parameters(self, src=0):
        if self.world > 1:
            for t in list(self.pipeline._fusion_network.parameters()) + list(self.pipeline._fusion_network.buffers()):
                dist.broadcast(t.data, src, group=self.group)

    def train_frame(self, batch, database, device, last=False):
        """Reference loop body (train_fusion.py:166-189) for one frame of this rank's scene stream."""
        bucket = self._flat_bucket()
        out = self.pipeline.fuse_training(batch, database, device)
        loss = self.criterion(out['tsdf_fused'], out['tsdf_target'])
        if loss.grad_fn:                                         # all rays masked -> constant, no backward
            loss.backward()
        if self.clipping:                                        # every iteration, on the running gradient
            torch.nn.utils.clip_grad_norm_(self.params, max_norm=1., norm_type=2)
        self.iteration += 1
        stepped = False
        if self.iteration % self.accumulation_steps == 0 or last:
            bucket = self._flat_bucket()                         # re-attach anything that replaced a gradient view
            if self.world > 1:
                ev = None
                if self.time_collective and bucket.is_cuda:
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record()
                dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group)
                if ev is not None:
                    ev[1].record()
                    self.collective_events.append(ev)
                bucket.div_(self.world)
            self.optimizer.step()
            bucket.zero_()                                       # == optimizer.zero_grad() with the views kept
            if self.scheduler is not None:
                self.scheduler.step()
            stepped = True
        return loss.detach(), stepped
