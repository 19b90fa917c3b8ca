"""Golden fixture for the network mirrors, produced from the REFERENCE's modules/model.py and
modules/adapnet.py (build container only):  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_nets.py"""
import hashlib
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def keys_digest(sd):
    return hashlib.sha256('\n'.join('%s %s' % (k, tuple(v.shape)) for k, v in sd.items()).encode()).hexdigest()


def main():
    _, _, _, ref_model, ref_adapnet = mg.ref_env()
    out = {}
    cfg = mg.make_config(48, 64)
    cfg.FUSION_MODEL.resx, cfg.FUSION_MODEL.resy = 64, 48
    for name, cls, sem in (('v3_sem', ref_model.FusionNet_v3, True), ('v3_nosem', ref_model.FusionNet_v3, False),
                           ('v2_nosem', ref_model.FusionNet_v2, False)):
        cfg.FUSION_MODEL.use_semantics = sem
        net = cls(cfg.FUSION_MODEL)
        out[name] = dict(n_keys=len(net.state_dict()), keys_sha256=keys_digest(net.state_dict()),
                         n_params=sum(p.numel() for p in net.parameters()))
    for stage in (1, 2):
        cfg.SEMANTIC_2D_MODEL.stage = stage
        net = ref_adapnet.AdapNet(cfg.SEMANTIC_2D_MODEL)
        sd = net.state_dict()
        out['adapnet_stage%d' % stage] = dict(
            n_keys=len(sd),
            keys_sorted_sha256=hashlib.sha256('\n'.join(sorted('%s %s' % (k, tuple(v.shape)) for k, v in sd.items())).encode()).hexdigest())
    seed, h, w = 77, 48, 64
    torch.manual_seed(seed)
    cfg.FUSION_MODEL.use_semantics = True
    net = ref_model.FusionNet_v3(cfg.FUSION_MODEL).eval()
    gen = torch.Generator().manual_seed(seed)
    x = {'tsdf_values': 0.05 * torch.randn(1, 9, h, w, generator=gen), 'tsdf_weights': torch.rand(1, 9, h, w, generator=gen),
         'tsdf_frame': 2 * torch.rand(1, 1, h, w, generator=gen), 'semantic_frame': torch.rand(1, 1, h, w, generator=gen)}
    with torch.no_grad():
        y = net(x)
    stride = 97
    out['v3_sem_forward'] = dict(seed=seed, h=h, w=w, sample_stride=stride,
                                 sample_values=[float(v) for v in y.reshape(-1)[::stride]])
    # AdapNet++ stage 2 (modules/adapnet.py:356-415): the reference's CPU logits for seeded parameters (filled in sorted
    # state_dict key order: identical values in the reference module and in the mirror) and a seeded input, with the
    # eval-time-active bottleneck dropout switched off (modules/adapnet.py:80-82)
    from online_joint_depthfusion_and_semantic_b200.synthetic import seeded_parameters
    cfg.SEMANTIC_2D_MODEL.stage = 2
    seg = seeded_parameters(ref_adapnet.AdapNet(cfg.SEMANTIC_2D_MODEL), 4321).eval()
    for m in seg.modules():
        if isinstance(m, ref_adapnet.BottleneckSSMA):
            m.dropout = False
    ah, aw = 32, 48
    gen = torch.Generator().manual_seed(99)
    m1, m2 = torch.randn(1, 3, ah, aw, generator=gen), torch.randn(1, 3, ah, aw, generator=gen)
    with torch.no_grad():
        res, aux1, aux2 = seg(m1, m2)
    astride = 53
    out['adapnet_stage2_forward'] = dict(param_seed=4321, input_seed=99, h=ah, w=aw, sample_stride=astride,
                                         logits_shape=list(res.shape), logits_abs_max=float(res.abs().max()),
                                         sample_values=[float(v) for v in res.reshape(-1)[::astride]],
                                         aux1_sample=[float(v) for v in aux1.reshape(-1)[::astride * 7]],
                                         aux2_sample=[float(v) for v in aux2.reshape(-1)[::astride * 7]])
    json.dump(out, open(os.path.join(HERE, 'nets.json'), 'w'), indent=1)
    print({k: (v if 'sample_values' not in v else '...') for k, v in out.items()})


if __name__ == '__main__':
    main()
