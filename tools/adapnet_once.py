#!/usr/bin/env python
"""A few eager forwards of the whole-network AdapNet++ engine at 240x320 (for `ncu` launch lists)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from online_joint_depthfusion_and_semantic_b200.config import fusion_config  # noqa: E402
from online_joint_depthfusion_and_semantic_b200.modules.adapnet import AdapNet  # noqa: E402

dev = torch.device('cuda:0')
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
net = AdapNet(fusion_config(240, 320).SEMANTIC_2D_MODEL).to(dev).eval()
net.set_bottleneck_dropout(False)
x1, x2 = torch.randn(1, 3, 240, 320, device=dev), torch.randn(1, 3, 240, 320, device=dev)
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
        net(x1, x2)
    torch.cuda.synchronize()
