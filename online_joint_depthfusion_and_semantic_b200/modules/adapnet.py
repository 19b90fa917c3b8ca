"""AdapNet++ / SSMA -- host-side mirror of the reference's modules/adapnet.py.

Same parameter tree and state_dict keys (1 091 keys in stage 2, including the unused
`res_n50_enc.fc`, e.g. `encoder_mod1.res_n50_enc.layer3.2.conv2a.weight`,
`decoder.stage3.8.weight`) so the reference's checkpoints load unchanged
(test_fusion.py:69-71):

  MultiScaleUnit = reference `BottleneckSSMA` (modules/adapnet.py:12-84): 1x1 -> {3x3 dil r1 | 3x3 dil r2}
                   -> concat -> 1x1, residual, ReLU, then the reference's eval-time-active Dropout(0.5)
  Encoder  (modules/adapnet.py:87-149): torchvision ResNet-50, conv1 re-made, layer2[3], layer3[2..5],
           layer4[0..2] replaced by multi-scale units, layer4 stride removed, two 24-channel skips
  eASPP    (modules/adapnet.py:152-216), Decoder (:219-317), SSMA (:320-354), AdapNet (:356-415)

Quirks kept on purpose: the bottleneck dropout is applied with a freshly built
`nn.Dropout`, i.e. it is active in eval mode (modules/adapnet.py:80-82) -- switch it off with
`AdapNet.set_bottleneck_dropout(False)` for deterministic runs; eASPP branch 5 and the decoder's
fuse-skip path skip their BatchNorm (modules/adapnet.py:204,312).  ImageNet weights are never
downloaded here (no network): the encoder starts from random init like any other checkpoint-loaded
module.
"""
import torch
from torch import nn
from torch.nn import functional as F
from torchvision.models import resnet50

from ._engine_cache import EngineOwner


class BottleneckSSMA(nn.Module):
    """Multi-scale residual unit.  (in_channels, out_channels, r1, r2, d3) as in the reference."""

    def __init__(self, in_channels, out_channels, r1, r2, d3, stride=1, downsample=None, copy_from=None, drop_out=True):
        super().__init__()
        self.dropout = drop_out
        self.dropout_default = drop_out
        half = d3 // 2
        self.conv2a = nn.Conv2d(out_channels, half, 3, stride=1, dilation=r1, padding=r1, bias=False)
        self.bn2a = nn.BatchNorm2d(half)
        self.conv2b = nn.Conv2d(out_channels, half, 3, stride=1, dilation=r2, padding=r2, bias=False)
        self.bn2b = nn.BatchNorm2d(half)
        self.conv3 = nn.Conv2d(d3, in_channels, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(in_channels)
        if copy_from is None:
            self.conv1 = nn.Conv2d(in_channels, out_channels, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(out_channels)
        else:                                   # reuse the torchvision unit's first 1x1 (modules/adapnet.py:41-46)
            self.conv1, self.bn1 = copy_from.conv1, copy_from.bn1
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = torch.cat((F.relu(self.bn2a(self.conv2a(y))), F.relu(self.bn2b(self.conv2b(y)))), dim=1)
        y = self.bn3(self.conv3(y))
        y = F.relu(y + (x if self.downsample is None else self.downsample(x)))
        if self.dropout:
            y = F.dropout(y, p=0.5, training=True)      # active in eval too, as in the reference
        return y


_LAYER3_UNITS = ((1024, 256, 1, 2, 256), (1024, 256, 1, 16, 256), (1024, 256, 1, 8, 256), (1024, 256, 1, 4, 256))
_LAYER4_UNITS = ((2048, 512, 2, 4, 512), (2048, 512, 2, 8, 512), (2048, 512, 2, 16, 512))


class Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.enc_skip2_conv = nn.Conv2d(256, 24, kernel_size=1, stride=1)
        self.enc_skip2_conv_bn = nn.BatchNorm2d(24)
        self.enc_skip1_conv = nn.Conv2d(512, 24, kernel_size=1, stride=1)
        self.enc_skip1_conv_bn = nn.BatchNorm2d(24)
        nn.init.kaiming_uniform_(self.enc_skip2_conv.weight, nonlinearity='relu')
        nn.init.kaiming_uniform_(self.enc_skip1_conv.weight, nonlinearity='relu')

        net = resnet50(weights=None)
        net.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        net.bn1 = nn.BatchNorm2d(64)
        net.layer2[-1] = BottleneckSSMA(512, 128, 1, 2, 64, copy_from=net.layer2[-1])
        for i, u in enumerate(_LAYER3_UNITS):
            net.layer3[i + 2] = BottleneckSSMA(*u, copy_from=net.layer3[i + 2], drop_out=(i == 0))
        for i, u in enumerate(_LAYER4_UNITS):
            down = None
            if i == 0:
                down = net.layer4[0].downsample
                down[0].stride = (1, 1)                 # output stride stays 16 (modules/adapnet.py:125-126)
            net.layer4[i] = BottleneckSSMA(*u, downsample=down, copy_from=net.layer4[i])
        self.res_n50_enc = net

    def forward_front(self, x):
        """Everything down to the first (strided) unit of layer3: returns (x at /16, skip2, skip1)."""
        n = self.res_n50_enc
        x = n.maxpool(n.relu(n.bn1(n.conv1(x))))
        x = n.layer1(x)
        s2 = self.enc_skip2_conv_bn(self.enc_skip2_conv(x))
        x = n.layer2(x)
        s1 = self.enc_skip1_conv_bn(self.enc_skip1_conv(x))
        return n.layer3[0](x), s2, s1

    def forward(self, x):
        n = self.res_n50_enc
        x, s2, s1 = self.forward_front(x)
        for unit in list(n.layer3)[1:]:
            x = unit(x)
        return n.layer4(x), s2, s1


def _aspp_branch(cin, mid, cout, rate):
    return nn.Sequential(
        nn.Conv2d(cin, mid, 1), nn.BatchNorm2d(mid), nn.ReLU(),
        nn.Conv2d(mid, mid, 3, dilation=rate, padding=rate), nn.BatchNorm2d(mid), nn.ReLU(),
        nn.Conv2d(mid, mid, 3, dilation=rate, padding=rate), nn.BatchNorm2d(mid), nn.ReLU(),
        nn.Conv2d(mid, cout, 1), nn.BatchNorm2d(cout), nn.ReLU())


class eASPP(nn.Module):
    def __init__(self, in_chs, mid_chs, out_chs):
        super().__init__()
        self.branch1_conv = nn.Conv2d(in_chs, out_chs, 1)
        self.branch1_bn = nn.BatchNorm2d(out_chs)
        self.branch234 = nn.ModuleList([_aspp_branch(in_chs, mid_chs, out_chs, r) for r in (3, 6, 12)])
        self.branch5_conv = nn.Conv2d(in_chs, out_chs, 1)
        self.branch5_bn = nn.BatchNorm2d(out_chs)       # present in checkpoints, unused (modules/adapnet.py:203-204)
        self.eASPP_fin_conv = nn.Conv2d(out_chs * 5, out_chs, 1)
        self.eASPP_fin_bn = nn.BatchNorm2d(out_chs)

    def forward(self, x):
        h, w = x.shape[2:]
        feats = [F.relu(self.branch1_bn(self.branch1_conv(x)))]
        feats += [b(x) for b in self.branch234]
        g = F.relu(self.branch5_conv(F.adaptive_avg_pool2d(x, 1)))
        feats.append(F.interpolate(g, size=(h, w), mode='bilinear', align_corners=True))
        return F.relu(self.eASPP_fin_bn(self.eASPP_fin_conv(torch.cat(feats, dim=1))))


class Decoder(nn.Module):
    def __init__(self, C, fusion=False):
        super().__init__()
        self.n_classes, self.fusion = C, fusion
        self.deconv1 = nn.ConvTranspose2d(256, 256, kernel_size=4, stride=2, padding=1)
        self.deconv1_bn = nn.BatchNorm2d(256)
        self.stage2 = nn.Sequential(
            nn.Conv2d(280, 256, 3, padding=1), nn.BatchNorm2d(256), nn.ReLU(),
            nn.Conv2d(256, 256, 3, padding=1), nn.BatchNorm2d(256), nn.ReLU(),
            nn.ConvTranspose2d(256, 256, kernel_size=4, stride=2, padding=1), nn.BatchNorm2d(256))
        self.stage3 = nn.Sequential(
            nn.Conv2d(280, 256, 3, padding=1), nn.BatchNorm2d(256), nn.ReLU(),
            nn.Conv2d(256, 256, 3, padding=1), nn.BatchNorm2d(256), nn.ReLU(),
            nn.Conv2d(256, C, 1), nn.BatchNorm2d(C),
            nn.ConvTranspose2d(C, C, kernel_size=8, stride=4, padding=2), nn.BatchNorm2d(C))
        self.aux_conv1 = nn.Conv2d(256, C, 1)
        self.aux_conv1_bn = nn.BatchNorm2d(C)
        self.aux_conv2 = nn.Conv2d(256, C, 1)
        self.aux_conv2_bn = nn.BatchNorm2d(C)
        self.fuse_conv1 = nn.Conv2d(256, 24, 1)
        self.fuse_conv1_bn = nn.BatchNorm2d(24)         # present in checkpoints, unused (modules/adapnet.py:311-312)
        self.fuse_conv2 = nn.Conv2d(256, 24, 1)
        self.fuse_conv2_bn = nn.BatchNorm2d(24)

    @staticmethod
    def _aux(x, conv, bn, scale):
        return F.interpolate(bn(conv(x)), scale_factor=scale, mode='bilinear', align_corners=True)

    def _join(self, x, skip, conv):
        if self.fusion:                                  # channel gate from the globally pooled decoder features
            skip = F.relu(conv(F.adaptive_avg_pool2d(x, 1))) * skip
        return torch.cat((x, skip), dim=1)

    def forward(self, x, skip1, skip2):
        x = F.relu(self.deconv1_bn(self.deconv1(x)))
        y1 = self._aux(x, self.aux_conv1, self.aux_conv1_bn, 8)
        x = self.stage2(self._join(x, skip1, self.fuse_conv1))
        y2 = self._aux(x, self.aux_conv2, self.aux_conv2_bn, 4)
        y3 = self.stage3(self._join(x, skip2, self.fuse_conv2))
        return y1, y2, y3


class SSMA(nn.Module):
    def __init__(self, features, bottleneck):
        super().__init__()
        reduced, doubled = int(features / bottleneck), 2 * features
        self.link = nn.Sequential(nn.Conv2d(doubled, reduced, 3, stride=1, padding=1), nn.ReLU(),
                                  nn.Conv2d(reduced, doubled, 3, stride=1, padding=1), nn.Sigmoid())
        self.final_conv = nn.Sequential(nn.Conv2d(doubled, features, 3, stride=1, padding=1), nn.BatchNorm2d(features))

    def forward(self, x1, x2):
        x = torch.cat((x1, x2), dim=1)
        return self.final_conv(x * self.link(x))


class AdapNet(EngineOwner, nn.Module):
    def __init__(self, config):
        super().__init__()
        self.stage = config.stage
        self.n_classes = config.n_classes
        self.fusion = self.stage != 1
        if self.stage == 1:
            self.encoder_mod1 = Encoder()
            self.eASPP = eASPP(2048, 64, 256)
        else:
            self.encoder_mod1 = Encoder()
            self.encoder_mod2 = Encoder()
            self.eASPP_mod1 = eASPP(2048, 64, 256)
            self.eASPP_mod2 = eASPP(2048, 64, 256)
            self.ssma_res = SSMA(256, 16)
            self.ssma_s1 = SSMA(24, 6)
            self.ssma_s2 = SSMA(24, 6)
        self.decoder = Decoder(self.n_classes, self.fusion)

    def no_resn50_dropout(self):
        """Reference helper (modules/adapnet.py:386-388): only layer3[2] of both encoders."""
        self._drop_engines()
        self.encoder_mod1.res_n50_enc.layer3[2].dropout = False
        self.encoder_mod2.res_n50_enc.layer3[2].dropout = False

    # ---- libojdf engine for the 15x20 tail (adapnet_engine.py); same invalidation rules as FusionNet
    _engine = None
    _full_engine = None
    use_engine = True
    whole_engine = True         # False: only the 15x20 tail runs on libojdf, the rest on the library
    aux_heads = True            # False: the engine returns [res, None, None] (the fusion pipeline reads only res)

    def _drop_engines(self):
        self._engine = None
        self._full_engine = None

    def engine_token(self):
        """Identity of the launch plans a CUDA graph captured over this module points into (Pipeline._segmentation)."""
        return (self._engine, self._full_engine)

    def train(self, mode=True):
        self._invalidate()
        return super().train(mode)

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._invalidate()
        return super().load_state_dict(*a, **k)

    def engine_ready(self, x):
        return self.use_engine and not self.training and not torch.is_grad_enabled() and x.is_cuda

    def _tail(self, pres):
        from .adapnet_engine import EncoderTailEngine
        h, w = pres[0].shape[-2:]
        self._hook_load_state_dict()
        self.engines_current()
        e = self._engine
        if e is None or (e.h, e.w) != (h, w) or e.device != pres[0].device:
            encs = [self.encoder_mod1] + ([self.encoder_mod2] if self.stage != 1 else [])
            heads = [self.eASPP] if self.stage == 1 else [self.eASPP_mod1, self.eASPP_mod2]
            e = self._engine = EncoderTailEngine(encs, heads, h, w, pres[0].device, flags=self.conv_flags)
        return e.forward(pres)

    def set_bottleneck_dropout(self, enabled):
        """Switch the eval-time-active dropout of every multi-scale unit (deterministic runs)."""
        self._drop_engines()
        for m in self.modules():
            if isinstance(m, BottleneckSSMA):
                m.dropout = bool(enabled) and m.dropout_default
        return self

    def _whole_engine(self, mod1):
        from .adapnet_engine import AdapNetEngine
        h, w = mod1.shape[-2:]
        self._hook_load_state_dict()
        self.engines_current()
        e = self._full_engine
        if e is None or (e.h, e.w) != (h, w) or e.device != mod1.device:
            e = self._full_engine = AdapNetEngine(self, h, w, mod1.device)
        return e

    def _whole(self, mod1, mod2):
        return self._whole_engine(mod1).forward(mod1, mod2)

    def whole_engine_ready(self, mod1):
        return self.engine_ready(mod1) and self.whole_engine and mod1.shape[0] == 1

    def segment(self, mod1, mod2=None):
        """(scores f32, ids u8, (1 + id) / n_classes f32), each (1,h,w): the per-pixel maximum / arg-max of
        softmax(main head) -- what the fusion pipeline consumes (modules/pipeline.py:57-58,184) -- computed by the
        launch plan plus one fused softmax-max kernel.  Only with whole_engine_ready(); forward() is the general path."""
        assert self.whole_engine_ready(mod1)
        return self._whole_engine(mod1).segment(mod1, mod2)

    def forward(self, mod1, mod2=None):
        if self.whole_engine_ready(mod1):
            return self._whole(mod1, mod2)
        if self.engine_ready(mod1):
            # front (conv1..layer3[0]) on the library, the 15x20 tail + eASPP on libojdf's kernels
            pre1, skip2, skip1 = self.encoder_mod1.forward_front(mod1)
            if self.stage == 1:
                x = self._tail([pre1])[0]
            else:
                pre2, m2_s2, m2_s1 = self.encoder_mod2.forward_front(mod2)
                x, x2 = self._tail([pre1, pre2])
                skip2 = self.ssma_s2(skip2, m2_s2)
                skip1 = self.ssma_s1(skip1, m2_s1)
                x = self.ssma_res(x, x2)
            aux1, aux2, res = self.decoder(x, skip1, skip2)
            return [res, aux1, aux2]
        if self.stage == 1:
            x, skip2, skip1 = self.encoder_mod1(mod1)
            x = self.eASPP(x)
        else:
            x, skip2, skip1 = self.encoder_mod1(mod1)
            x2, m2_s2, m2_s1 = self.encoder_mod2(mod2)
            x, x2 = self.eASPP_mod1(x), self.eASPP_mod2(x2)
            skip2 = self.ssma_s2(skip2, m2_s2)
            skip1 = self.ssma_s1(skip1, m2_s1)
            x = self.ssma_res(x, x2)
        aux1, aux2, res = self.decoder(x, skip1, skip2)
        return [res, aux1, aux2]
