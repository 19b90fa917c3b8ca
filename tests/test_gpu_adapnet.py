"""AdapNet++ on libojdf's kernels (whole-network engine, and the tail-only engine) vs the plain PyTorch fp32
forward of the same module (bit-identical to the reference's modules/adapnet.py on CPU, tests/test_networks_cpu.py).
Tolerance: the north star's 1e-4 relative on the semantic logits (max |a-b| <= 1e-4 * max |b|)."""
import pytest
import torch

from online_joint_depthfusion_and_semantic_b200.config import fusion_config
from online_joint_depthfusion_and_semantic_b200.modules.adapnet import AdapNet

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def _net(stage, seed=11):
    torch.manual_seed(seed)
    cfg = fusion_config(240, 320)
    cfg.SEMANTIC_2D_MODEL.stage = stage
    net = AdapNet(cfg.SEMANTIC_2D_MODEL)
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(0.6 + 0.8 * torch.rand(m.num_features, generator=g))
    net.set_bottleneck_dropout(False)
    return net.to(DEV).eval()


@pytest.mark.parametrize('stage,h,w,whole', [(2, 240, 320, True), (1, 64, 96, True), (2, 48, 64, True),
                                              (2, 240, 320, False), (1, 64, 96, False)])
def test_engine_matches_torch_fp32(stage, h, w, whole):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = _net(stage)
    net.whole_engine = whole
    g = torch.Generator().manual_seed(3)
    x1, x2 = torch.randn(1, 3, h, w, generator=g).to(DEV), torch.randn(1, 3, h, w, generator=g).to(DEV)
    args = (x1, x2) if stage == 2 else (x1,)
    with torch.no_grad():
        net.use_engine = False
        ref = net(*args)
        net.use_engine = True
        out = net(*args)
        torch.cuda.synchronize()
    assert (net._full_engine if whole else net._engine) is not None
    for a, b in zip(out, ref):
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= 1e-4 * scale, (float((a - b).abs().max()), scale)
    # label agreement of the arg-max (what the pipeline integrates)
    assert float((out[0].argmax(1) == ref[0].argmax(1)).float().mean()) > 0.999


def test_dropout_quirk_stays_random_with_engine():
    net = _net(2)
    net.set_bottleneck_dropout(True)
    x1, x2 = torch.randn(1, 3, 48, 64, device=DEV), torch.randn(1, 3, 48, 64, device=DEV)
    with torch.no_grad():
        a, b = net(x1, x2)[0].clone(), net(x1, x2)[0].clone()
    assert float((a - b).abs().max()) > 0          # eval-time-active dropout, as in the reference


@pytest.mark.parametrize('stage,h,w', [(2, 240, 320), (1, 64, 96)])
def test_segment_scores_ids_vs_torch_softmax_max(stage, h, w):
    """Row a3: `AdapNet.segment` (launch plan + own stem kernel + fused softmax / max / arg-max kernel) against
    torch.softmax(logits, 1).max(1) of the plain fp32 forward (modules/pipeline.py:57-58,184)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = _net(stage)
    g = torch.Generator().manual_seed(5)
    x1, x2 = torch.randn(1, 3, h, w, generator=g).to(DEV), torch.randn(1, 3, h, w, generator=g).to(DEV)
    args = (x1, x2) if stage == 2 else (x1,)
    with torch.no_grad():
        net.use_engine = False
        probs = torch.softmax(net(*args)[0], dim=1)
        ref_scores, ref_ids = probs.max(dim=1)
        net.use_engine = True
        net.aux_heads = False
        scores, ids, frame = net.segment(*args)
        torch.cuda.synchronize()
    assert scores.shape == (1, h, w) and ids.dtype == torch.uint8
    assert float((ids.long() == ref_ids).float().mean()) > 0.999
    same = ids.long() == ref_ids
    assert float((scores[same] - ref_scores[same]).abs().max()) <= 1e-4 * float(ref_scores.max())
    assert torch.allclose(frame, (1 + ids.float()) / net.decoder.n_classes, rtol=1e-6, atol=0)
