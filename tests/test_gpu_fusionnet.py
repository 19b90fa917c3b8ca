"""FusionNet on libojdf's fused fp32 kernels vs the plain PyTorch fp32 forward of the same module
(which is bit-identical to the reference's modules/model.py on CPU, see tests/test_networks_cpu.py).
Tolerance (written here as the north star states it): 1e-4 relative -- max |a-b| <= 1e-4 * max |b|."""
import pytest
import torch

from online_joint_depthfusion_and_semantic_b200.config import fusion_config
from online_joint_depthfusion_and_semantic_b200.modules.model import FusionNet_v2, FusionNet_v3

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def _net(cls, h, w, use_sem, seed=3):
    torch.manual_seed(seed)
    cfg = fusion_config(h, w, use_semantics=use_sem)
    cfg.FUSION_MODEL.resx, cfg.FUSION_MODEL.resy = w, h
    net = cls(cfg.FUSION_MODEL)
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(0.2 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.num_features, generator=g))
    return net.to(DEV).eval()


def _inputs(h, w, seed=5):
    g = torch.Generator().manual_seed(seed)
    return {'tsdf_values': (0.05 * torch.randn(1, 9, h, w, generator=g)).to(DEV),
            'tsdf_weights': (10 * torch.rand(1, 9, h, w, generator=g)).to(DEV),
            'tsdf_frame': (0.5 + 2 * torch.rand(1, 1, h, w, generator=g)).to(DEV),
            'semantic_frame': ((1 + torch.randint(0, 30, (1, 1, h, w), generator=g)).float() / 30).to(DEV)}


@pytest.mark.parametrize('cls,use_sem,h,w', [(FusionNet_v3, True, 48, 64), (FusionNet_v3, False, 48, 64),
                                             (FusionNet_v2, False, 40, 56), (FusionNet_v3, True, 240, 320),
                                             (FusionNet_v3, True, 37, 53)])
def test_engine_matches_torch_fp32(cls, use_sem, h, w):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = _net(cls, h, w, use_sem)
    x = _inputs(h, w)
    with torch.no_grad():
        net.use_engine = False
        ref = net(x)
        net.use_engine = True
        assert net.engine_ready(x['tsdf_values'])
        out = net(x)
        torch.cuda.synchronize()
    assert out.shape == ref.shape == (1, 9, h, w)
    scale = float(ref.abs().max())
    assert scale > 1e-3
    err = float((out - ref).abs().max())
    assert err <= 1e-4 * scale, (err, scale)
    # and against the fp32 CPU forward (no cuDNN involved at all)
    cpu = net.cpu()
    with torch.no_grad():
        cpu.use_engine = False
        ref_cpu = cpu({k: v.cpu() for k, v in x.items()})
    assert float((out.cpu() - ref_cpu).abs().max()) <= 1e-4 * float(ref_cpu.abs().max())


def test_engine_dropped_when_parameters_change():
    net = _net(FusionNet_v3, 32, 32, True)
    x = _inputs(32, 32)
    with torch.no_grad():
        a = net(x).clone()
        assert net._engine is not None
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        sd['pred.4.pred.6.bias'] += 0.05
        net.load_state_dict(sd)
        assert net._engine is None
        b = net(x)
        net.use_engine = False
        ref = net(x)
    assert float((a - b).abs().max()) > 1e-3
    assert float((b - ref).abs().max()) <= 1e-4 * float(ref.abs().max())


def test_training_mode_uses_autograd_path():
    net = _net(FusionNet_v3, 32, 32, True).train()
    x = _inputs(32, 32)
    assert not net.engine_ready(x['tsdf_values'])
    y = net(x)
    y.sum().backward()
    assert net.block0[0].block[0].weight.grad is not None
