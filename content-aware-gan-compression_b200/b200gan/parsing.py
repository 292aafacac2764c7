"""Inference form of the face parser that supplies the content mask of the KD loss (second half).

The reference builds `BiSeNet(n_classes=19)` (Util/face_parsing/BiSeNet.py:230-254, ResNet18 context path :96-139 /
Util/face_parsing/resnet.py:57-83), loads 79999_iter.pth, calls `.eval()` and only ever reads `net(x)[0]`
(Util/content_aware_pruning.py:18-36, :52, :85).  `FaceParser.from_state_dict` takes that state_dict and keeps what
inference of output 0 needs:

  * every BatchNorm folded into the convolution in front of it (eval mode: a per-channel affine);
  * the two auxiliary heads (`conv_out16`, `conv_out32`) and their 19-channel 512^2 upsamples -- training-time deep
    supervision whose results the caller drops -- are not evaluated: outputs 1 and 2 of the tuple are None.

The convolutions are LIBRARY calls (`F.conv2d`: cuDNN, TF32 where torch allows it, channels-last): a third-party network
outside the generator path; the package's own kernels take over on both sides of it (b200gan/maskglue.py).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F
from torch import nn


def _fold(sd: Dict[str, torch.Tensor], conv: str, bn: str, eps: float = 1e-5):
    """conv (no bias) + BatchNorm2d in eval mode -> (weight, bias)."""
    w = sd[conv + '.weight'].detach().double()
    g, b = sd[bn + '.weight'].detach().double(), sd[bn + '.bias'].detach().double()
    m, v = sd[bn + '.running_mean'].detach().double(), sd[bn + '.running_var'].detach().double()
    s = g / torch.sqrt(v + eps)
    return w * s.view(-1, 1, 1, 1), b - m * s


class FaceParser(nn.Module):
    N_LAYERS = (('layer1', 64, 1), ('layer2', 128, 2), ('layer3', 256, 2), ('layer4', 512, 2))

    def __init__(self, folded: Dict[str, torch.Tensor]):
        super().__init__()
        for k, t in folded.items():
            t = t.float().contiguous()
            if t.ndim == 4:
                t = t.contiguous(memory_format=torch.channels_last)
            self.register_buffer(k.replace('.', '_'), t)

    def _p(self, key):
        return getattr(self, key.replace('.', '_') + '_w'), getattr(self, key.replace('.', '_') + '_b', None)

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor]) -> 'FaceParser':
        """sd: state_dict of the reference BiSeNet (keys `cp.resnet.*`, `cp.arm16.*`, ..., `ffm.*`, `conv_out.*`)."""
        sd = {k[7:] if k.startswith('module.') else k: v for k, v in sd.items()}
        f = {}

        def put(key, conv, bn=None):
            if bn is None:
                f[key + '_w'] = sd[conv + '.weight'].detach().double()
            else:
                f[key + '_w'], f[key + '_b'] = _fold(sd, conv, bn)
        r = 'cp.resnet.'
        put('stem', r + 'conv1', r + 'bn1')
        cin = 64
        for layer, cout, stride in cls.N_LAYERS:
            for blk in (0, 1):
                p = f'{r}{layer}.{blk}.'
                put(f'{layer}.{blk}.c1', p + 'conv1', p + 'bn1')
                put(f'{layer}.{blk}.c2', p + 'conv2', p + 'bn2')
                if blk == 0 and (cin != cout or stride != 1):
                    put(f'{layer}.{blk}.ds', p + 'downsample.0', p + 'downsample.1')
            cin = cout
        for arm in ('arm16', 'arm32'):
            put(f'{arm}.conv', f'cp.{arm}.conv.conv', f'cp.{arm}.conv.bn')
            put(f'{arm}.atten', f'cp.{arm}.conv_atten', f'cp.{arm}.bn_atten')
        for name in ('conv_head32', 'conv_head16', 'conv_avg'):
            put(name, f'cp.{name}.conv', f'cp.{name}.bn')
        put('ffm.convblk', 'ffm.convblk.conv', 'ffm.convblk.bn')
        put('ffm.conv1', 'ffm.conv1')
        put('ffm.conv2', 'ffm.conv2')
        put('out.conv', 'conv_out.conv.conv', 'conv_out.conv.bn')
        put('out.cls', 'conv_out.conv_out')
        return cls(f)

    fused = True        # CUDA: cuDNN's fused conv + bias (+ residual) + ReLU calls instead of conv -> add -> relu launches

    def _conv(self, x, key, stride=1, padding=1, relu=True, add=None):
        """[relu](conv(x, W) + b [+ add])."""
        w, b = self._p(key)
        w = w.to(x.dtype)
        b = None if b is None else b.to(x.dtype)
        if self.fused and x.is_cuda and relu and b is not None and x.dtype == torch.float32 and x.shape[2] * x.shape[3] > 1:
            s2, p2, d2 = (stride, stride), (padding, padding), (1, 1)
            if add is None:
                return torch.cudnn_convolution_relu(x, w, b, s2, p2, d2, 1)
            return torch.cudnn_convolution_add_relu(x, w, add, 1.0, b, s2, p2, d2, 1)
        y = F.conv2d(x, w, b, stride, padding)
        if add is not None:
            y = y + add
        return F.relu(y) if relu else y

    def _arm(self, x, arm):                                   # AttentionRefinementModule, BiSeNet.py:76-93
        feat = self._conv(x, f'{arm}.conv')
        atten = torch.sigmoid(self._conv(feat.mean((2, 3), keepdim=True), f'{arm}.atten', padding=0, relu=False))
        return feat * atten

    def forward(self, x, lowres=False):
        h, w = x.shape[2:]
        if x.is_cuda:
            x = x.contiguous(memory_format=torch.channels_last)      # cuDNN's tensor-core kernels are NHWC
        # ResNet18 trunk (resnet.py:69-78)
        t = F.max_pool2d(self._conv(x, 'stem', stride=2, padding=3), 3, 2, 1)
        feats = []
        for layer, _, stride in self.N_LAYERS:
            for blk in (0, 1):
                s = stride if blk == 0 else 1
                short = t
                if hasattr(self, f'{layer}_{blk}_ds_w'):
                    short = self._conv(t, f'{layer}.{blk}.ds', stride=s, padding=0, relu=False)
                t = self._conv(self._conv(t, f'{layer}.{blk}.c1', stride=s), f'{layer}.{blk}.c2', add=short)
            feats.append(t)
        feat8, feat16, feat32 = feats[1], feats[2], feats[3]
        # context path (BiSeNet.py:112-133); nearest upsampling of a 1x1 map is a broadcast
        avg = self._conv(feat32.mean((2, 3), keepdim=True), 'conv_avg', padding=0)
        f32 = self._arm(feat32, 'arm32') + avg
        f32 = self._conv(F.interpolate(f32, feat16.shape[2:], mode='nearest'), 'conv_head32')
        f16 = self._arm(feat16, 'arm16') + f32
        f16 = self._conv(F.interpolate(f16, feat8.shape[2:], mode='nearest'), 'conv_head16')
        # feature fusion (BiSeNet.py:190-200) and the main head (:47-50), upsampled as :247
        feat = self._conv(torch.cat([feat8, f16], 1), 'ffm.convblk', padding=0)
        atten = self._conv(feat.mean((2, 3), keepdim=True), 'ffm.conv1', padding=0)
        atten = torch.sigmoid(self._conv(atten, 'ffm.conv2', padding=0, relu=False))
        feat = feat * atten + feat
        out = self._conv(self._conv(feat, 'out.conv'), 'out.cls', padding=0, relu=False)
        if lowres:
            return out
        out = F.interpolate(out, (h, w), mode='bilinear', align_corners=True)
        return out, None, None

    def scores_lowres(self, x):
        """Class scores BEFORE the final bilinear upsample (1/8 resolution): b200gan.maskglue evaluates the upsample
        (BiSeNet.py:247, align_corners=True) inside its mask kernel instead of materialising [N,19,512,512]."""
        return self.forward(x, lowres=True)


def synthetic_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """A random BiSeNet-shaped state_dict (the 53 MB 79999_iter.pth is not redistributable offline): He-normal
    convolutions, BatchNorm statistics around (0, 1).  For timing runs and tests."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cout, cin, k):
        sd[name + '.weight'] = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5

    def bn(name, c):
        sd[name + '.weight'] = 1 + 0.1 * torch.randn(c, generator=g)
        sd[name + '.bias'] = 0.1 * torch.randn(c, generator=g)
        sd[name + '.running_mean'] = 0.1 * torch.randn(c, generator=g)
        sd[name + '.running_var'] = 1 + 0.1 * torch.rand(c, generator=g)
    r = 'cp.resnet.'
    conv(r + 'conv1', 64, 3, 7)
    bn(r + 'bn1', 64)
    cin = 64
    for layer, cout, stride in FaceParser.N_LAYERS:
        for blk in (0, 1):
            p = f'{r}{layer}.{blk}.'
            conv(p + 'conv1', cout, cin if blk == 0 else cout, 3)
            bn(p + 'bn1', cout)
            conv(p + 'conv2', cout, cout, 3)
            bn(p + 'bn2', cout)
            if blk == 0 and (cin != cout or stride != 1):
                conv(p + 'downsample.0', cout, cin, 1)
                bn(p + 'downsample.1', cout)
        cin = cout
    for arm, c in (('arm16', 256), ('arm32', 512)):
        conv(f'cp.{arm}.conv.conv', 128, c, 3)
        bn(f'cp.{arm}.conv.bn', 128)
        conv(f'cp.{arm}.conv_atten', 128, 128, 1)
        bn(f'cp.{arm}.bn_atten', 128)
    for name, c, k in (('conv_head32', 128, 3), ('conv_head16', 128, 3), ('conv_avg', 512, 1)):
        conv(f'cp.{name}.conv', 128, c, k)
        bn(f'cp.{name}.bn', 128)
    conv('ffm.convblk.conv', 256, 256, 1)
    bn('ffm.convblk.bn', 256)
    conv('ffm.conv1', 64, 256, 1)
    conv('ffm.conv2', 256, 64, 1)
    conv('conv_out.conv.conv', 256, 256, 3)
    bn('conv_out.conv.bn', 256)
    conv('conv_out.conv_out', 19, 256, 1)
    return sd
