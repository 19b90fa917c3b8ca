#!/usr/bin/env python
"""Where one online-training frame (bench.py --mode train: Pipeline.fuse_training + FusionLoss + backward) spends its time."""
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from online_joint_depthfusion_and_semantic_b200.training import FusionLoss, PolynomialLR, ShardedFusionTrainer  # noqa: E402

dev = torch.device('cuda:0')
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
cfg, pipe, db, host_frames = bench.build_world(dev, 0, scenes_per_rank=2, frames=8)
pipe.train()
pipe._semantic_2d_network.eval()
for p_ in pipe._semantic_2d_network.parameters():
    p_.requires_grad_(False)
opt = torch.optim.RMSprop(pipe._fusion_network.parameters(), lr=1e-5, momentum=0.9, weight_decay=0.01, eps=1e-9)
trainer = ShardedFusionTrainer(pipe, opt, PolynomialLR(opt, max_iter=50000), FusionLoss(), accumulation_steps=8, clipping=True)
frames = [bench.to_device_frame(hb, dev) for hb in host_frames]


def step(i):
    b = dict(frames[i % len(frames)])
    b['tof_depth'] = b['tof_depth'].clone()
    return trainer.train_frame(b, db, dev)


for i in range(8):
    step(i)
torch.cuda.synchronize()
t0 = time.time()
for i in range(8):
    step(i)
torch.cuda.synchronize()
print('wall per frame %.1f ms' % ((time.time() - t0) / 8 * 1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(4):
        step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=30, max_name_column_width=70))
print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=15, max_name_column_width=70))
