#!/usr/bin/env python
"""Where AdapNet++ stage-2 spends its GPU time at 240x320 (torch.profiler, per conv shape)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from online_joint_depthfusion_and_semantic_b200.config import fusion_config  # noqa: E402
from online_joint_depthfusion_and_semantic_b200.modules.adapnet import AdapNet  # noqa: E402

dev = torch.device('cuda:0')
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
net = AdapNet(fusion_config(240, 320).SEMANTIC_2D_MODEL).to(dev).eval()
x1, x2 = torch.randn(1, 3, 240, 320, device=dev), torch.randn(1, 3, 240, 320, device=dev)
with torch.no_grad():
    for _ in range(3):
        net(x1, x2)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        for _ in range(5):
            net(x1, x2)
        torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=28, max_shapes_column_width=70))
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=60))
