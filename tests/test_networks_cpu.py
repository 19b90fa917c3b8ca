"""Host mirror of the two networks: parameter trees and (CPU, fp32) outputs must equal the
reference's.  The reference itself is not available outside the build container, so the pin is a
fixture produced from it (tests/golden/make_golden_nets.py): the state_dict key lists and output
digests for seeded weights and inputs."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from online_joint_depthfusion_and_semantic_b200.config import fusion_config
from online_joint_depthfusion_and_semantic_b200.modules.adapnet import AdapNet
from online_joint_depthfusion_and_semantic_b200.modules.model import FusionNet_v2, FusionNet_v3

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'nets.json')


@pytest.fixture(scope='module')
def gold():
    return json.load(open(GOLD))


def _keys_digest(sd):
    return hashlib.sha256('\n'.join('%s %s' % (k, tuple(v.shape)) for k, v in sd.items()).encode()).hexdigest()


def test_fusionnet_state_dict_matches_reference(gold):
    cfg = fusion_config(48, 64)
    cfg.FUSION_MODEL.resx, cfg.FUSION_MODEL.resy = 64, 48
    for name, cls, sem in (('v3_sem', FusionNet_v3, True), ('v3_nosem', FusionNet_v3, False), ('v2_nosem', FusionNet_v2, False)):
        cfg.FUSION_MODEL.use_semantics = sem
        net = cls(cfg.FUSION_MODEL)
        sd = net.state_dict()
        assert len(sd) == gold[name]['n_keys']
        assert _keys_digest(sd) == gold[name]['keys_sha256']
        assert sum(p.numel() for p in net.parameters()) == gold[name]['n_params']


def test_adapnet_state_dict_matches_reference(gold):
    cfg = fusion_config(48, 64)
    for stage in (1, 2):
        cfg.SEMANTIC_2D_MODEL.stage = stage
        net = AdapNet(cfg.SEMANTIC_2D_MODEL)
        sd = net.state_dict()
        g = gold['adapnet_stage%d' % stage]
        assert len(sd) == g['n_keys']
        assert hashlib.sha256('\n'.join(sorted('%s %s' % (k, tuple(v.shape)) for k, v in sd.items())).encode()).hexdigest() == g['keys_sorted_sha256']


def test_fusionnet_output_matches_reference_values(gold):
    """Same seeded weights + inputs as make_golden_nets.py; the reference's CPU output is stored."""
    g = gold['v3_sem_forward']
    torch.manual_seed(g['seed'])
    cfg = fusion_config(g['h'], g['w'])
    cfg.FUSION_MODEL.resx, cfg.FUSION_MODEL.resy = g['w'], g['h']
    net = FusionNet_v3(cfg.FUSION_MODEL).eval()
    gen = torch.Generator().manual_seed(g['seed'])
    x = {'tsdf_values': 0.05 * torch.randn(1, 9, g['h'], g['w'], generator=gen),
         'tsdf_weights': torch.rand(1, 9, g['h'], g['w'], generator=gen),
         'tsdf_frame': 2 * torch.rand(1, 1, g['h'], g['w'], generator=gen),
         'semantic_frame': torch.rand(1, 1, g['h'], g['w'], generator=gen)}
    with torch.no_grad():
        y = net(x)
    ref = np.asarray(g['sample_values'], dtype=np.float32)
    got = y.reshape(-1)[::g['sample_stride']].numpy()
    assert got.shape == ref.shape
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-7)


def test_adapnet_output_matches_reference_values(gold):
    """AdapNet++ stage 2 mirror vs the REFERENCE's own CPU logits (tests/golden/make_golden_nets.py): same seeded
    parameters (filled in sorted state_dict key order) and input, bottleneck dropout off on both sides."""
    from online_joint_depthfusion_and_semantic_b200.synthetic import seeded_parameters
    g = gold['adapnet_stage2_forward']
    cfg = fusion_config(g['h'], g['w'])
    cfg.SEMANTIC_2D_MODEL.stage = 2
    net = seeded_parameters(AdapNet(cfg.SEMANTIC_2D_MODEL), g['param_seed']).eval()
    net.set_bottleneck_dropout(False)
    gen = torch.Generator().manual_seed(g['input_seed'])
    m1, m2 = torch.randn(1, 3, g['h'], g['w'], generator=gen), torch.randn(1, 3, g['h'], g['w'], generator=gen)
    with torch.no_grad():
        res, aux1, aux2 = net(m1, m2)
    assert list(res.shape) == g['logits_shape']
    scale = g['logits_abs_max']
    for got, key, stride in ((res, 'sample_values', g['sample_stride']), (aux1, 'aux1_sample', g['sample_stride'] * 7),
                             (aux2, 'aux2_sample', g['sample_stride'] * 7)):
        ref = np.asarray(g[key], dtype=np.float32)
        out = got.reshape(-1)[::stride].numpy()
        assert out.shape == ref.shape
        assert float(np.abs(out - ref).max()) <= 1e-5 * scale, key
