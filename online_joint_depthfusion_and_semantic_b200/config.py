"""Attribute-dict configuration with the keys the hot path reads:
FUSION_MODEL.{name,n_points,n_tail_points,growth_factor,use_semantics,output_scale},
SEMANTIC_2D_MODEL.{stage,n_classes}, DATA.{semantics,semantic_strategy,input,resx,resy,init_value},
SETTINGS.{gpu,device,implementation}.  Stands in for EasyDict (utils/loading.py:9-19), which is
not installed here; raises AttributeError on missing keys like EasyDict does."""
import torch


class Config(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        for key, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, Config):
                self[key] = Config(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def fusion_config(h=240, w=320, semantics='class30', semantic_strategy='predict', use_semantics=True,
                  n_classes=30, stage=2, device='cuda:0', input_key='tof_depth', init_value=0.1):
    """configs/fusion/replica_accuracy.yaml with the frame size of BASELINE.json."""
    return Config(
        SETTINGS=dict(gpu=True, device=torch.device(device), implementation='efficient', seed=1911),
        FUSION_MODEL=dict(name='v3', output_scale=1.0, n_points=9, n_tail_points=7, growth_factor=6,
                          use_semantics=bool(use_semantics)),
        SEMANTIC_2D_MODEL=dict(stage=stage, n_classes=n_classes),
        DATA=dict(semantics=semantics, semantic_strategy=semantic_strategy, semantic_grid=True, input=input_key,
                  resx=w, resy=h, init_value=init_value),
    )
