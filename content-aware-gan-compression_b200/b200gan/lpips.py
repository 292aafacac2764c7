"""LPIPS-VGG16 perceptual loss of the KD step on the package's own kernels.

Reference: train.py:172-182 calls `lpips.PerceptualLoss(model='net-lin', net='vgg')` (lpips/__init__.py:13-41), i.e.
`PNetLin.forward(target, pred)` (lpips/networks_basic.py:62-92) over torchvision's VGG16 `features[0:30]` cut into five
slices (lpips/pretrained_networks.py:97-137), everything frozen: the student only needs the DATA gradient towards `pred`.

    ScalingLayer -> conv1_1 .. relu5_3 (13 conv3x3 + ReLU, 4 MaxPool2d)      both images as ONE batch of 2N
    per tap (relu1_2, relu2_2, relu3_3, relu4_3, relu5_3):
        unit-normalise over channels, squared difference, 1x1 `lin` layer, spatial mean; sum over the five taps

Here the whole distance is one autograd node on NHWC-p buffers:

    forward   conv1_1: cagc_rgb_conv3x3_fwd (scaling layer on the operand load, bias + ReLU on the store)
              12 x cagc_conv2d_ws with a ReLU epilogue (tcgen05 TF32 implicit GEMM, or the exact-fp32 engine)
              4 x cagc_maxpool2_nhwc, 5 x cagc_lpips_head_fwd
    backward  5 x cagc_lpips_head_bwd; per layer ONE data-gradient convolution; the ReLU mask of a conv -> ReLU -> conv chain
              is that convolution's epilogue (cagc_conv2d_mask_ws), tapped / pooled layers take ONE masking pass
              (cagc_relu_pool_bwd: max-pool backward + tap gradient + ReLU mask); conv1_1: cagc_rgb_conv3x3_bwd straight
              to the image gradient

The dropout in front of every `lin` layer is the identity: DistModel.initialize puts the net in eval mode
(lpips/dist_model.py:98-99).  No weight gradient exists on this path (requires_grad=False everywhere in the reference).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ._lib import lib, check, stream_of, require_cuda, conv_workspace, ptr
from . import config
from .dconv import _prep, _empty
from .modconv import _timed

# torchvision.models.vgg16().features[0:30]: output channels per conv, 'M' = MaxPool2d(2, 2) (pretrained_networks.py:100-114)
VGG16_CFG = (64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512)
TAPS = (1, 3, 6, 9, 12)                 # conv indices whose ReLU output is tapped (relu1_2 .. relu5_3)
POOL_BEFORE = (2, 4, 7, 10)             # conv indices whose input went through a max-pool
SHIFT = (-.030, -.088, -.188)           # ScalingLayer, networks_basic.py:94-101
SCALE = (.458, .448, .450)


# development knob: ReLU masks of the conv -> ReLU -> conv chains as data-gradient epilogues (0: one masking pass per layer)
_MASK_EPILOGUE = os.environ.get('CAGC_LPIPS_MASK_EPILOGUE', '1') != '0'


def _f3(vals):
    return (C.c_float * 3)(*[float(v) for v in vals])


class _LpipsFn(Function):
    @staticmethod
    def forward(ctx, pred, target, mod, algo):
        require_cuda(pred, 'LPIPS')
        require_cuda(target, 'LPIPS')
        if pred.shape != target.shape or pred.ndim != 4 or pred.shape[1] != 3:
            raise RuntimeError(f'LPIPS: two [N,3,H,W] images expected, got {tuple(pred.shape)} and {tuple(target.shape)}')
        n, _, h, w = pred.shape
        if h % 16 or w % 16:
            raise RuntimeError(f'LPIPS: image sides must be multiples of 16 (four 2x2 max-pools), got {h}x{w}')
        dev = pred.device
        ws, bs = list(mod.conv_weights), list(mod.conv_biases)
        acts: List[torch.Tensor] = []
        val = torch.empty(n, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            st = stream_of(pred)
            # conv1_1 on both images: rows [0, n) = pred (the student), rows [n, 2n) = target (the teacher)
            c0 = ws[0].shape[0]
            a = _empty(2 * n, h, w, c0, dev)
            for half, img in ((0, pred.detach()), (1, target.detach())):
                sb, sc, sh, sw = img.stride()
                _timed('lpips_rgb_conv', 2.0 * n * h * w * 27 * c0, 4.0 * n * h * w * (3 + c0),
                       lambda: check(lib.cagc_rgb_conv3x3_fwd(st, img.data_ptr(), sb, sc, sh, sw, ws[0].data_ptr(),
                                                              bs[0].data_ptr(), mod._shift, mod._scale,
                                                              a[half * n:].data_ptr(), n, h, w, c0, c0), 'lpips.conv1_1'))
            acts.append(a)
            x, hh, ww = a, h, w
            for i in range(1, 13):
                cin, cout = ws[i].shape[1], ws[i].shape[0]
                if i in POOL_BEFORE:
                    xp = _empty(2 * n, hh // 2, ww // 2, cin, dev)
                    check(lib.cagc_maxpool2_nhwc(st, x.data_ptr(), xp.data_ptr(), 2 * n, hh, ww, cin), 'lpips.maxpool')
                    x, hh, ww = xp, hh // 2, ww // 2
                p, tcf, _ = _prep(ws[i], bs[i], 1.0, False, algo)
                y = _empty(2 * n, hh, ww, cout, dev)
                _vconv(st, x, p.w_fwd, p.bias_p, y, 2 * n, hh, ww, cin, cout, True, tcf, f'lpips.conv{i}')
                acts.append(y)
                x = y
            for kk, i in enumerate(TAPS):
                f = acts[i]
                cch, hw = f.shape[3], f.shape[1] * f.shape[2]
                nblk = int(lib.cagc_lpips_head_blocks(n, hw))
                partial = torch.empty(n * nblk, device=dev, dtype=torch.float32)
                check(lib.cagc_lpips_head_fwd(st, f.data_ptr(), f[n:].data_ptr(), mod.lin_weights[kk].data_ptr(),
                                              partial.data_ptr(), val.data_ptr(), n, hw, cch, int(kk > 0)), 'lpips.head')
        ctx.mod, ctx.algo, ctx.n = mod, algo, n
        ctx.acts = acts                      # intermediates of a frozen network: nothing autograd needs to version-check
        return val.view(n, 1, 1, 1)

    @staticmethod
    @once_differentiable
    def backward(ctx, gval):
        mod, algo, n, acts = ctx.mod, ctx.algo, ctx.n, ctx.acts
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        dev = gval.device
        ws, bs = list(mod.conv_weights), list(mod.conv_biases)
        gv = gval.reshape(n).contiguous()
        with torch.cuda.device(dev):
            st = stream_of(gval)
            g = None        # gradient w.r.t. the OUTPUT of conv i (student half), before its ReLU mask, or ...
            gz = None       # ... the masked gradient itself, when the data-gradient convolution above applied the mask
            for i in range(12, 0, -1):
                a = acts[i]
                _, hh, ww, cout = a.shape
                cin = ws[i].shape[1]
                if gz is None:
                    gfeat = None
                    if i in TAPS:
                        kk = TAPS.index(i)
                        gfeat = _empty(n, hh, ww, cout, dev)
                        check(lib.cagc_lpips_head_bwd(st, a.data_ptr(), a[n:].data_ptr(), mod.lin_weights[kk].data_ptr(),
                                                      gv.data_ptr(), gfeat.data_ptr(), n, hh * ww, cout), 'lpips.head^T')
                    gz = _empty(n, hh, ww, cout, dev)
                    if i + 1 in POOL_BEFORE:     # the output also feeds a max-pool: g lives on the pooled grid
                        check(lib.cagc_relu_pool_bwd(st, a.data_ptr(), g.data_ptr(), ptr(gfeat), gz.data_ptr(), n, hh, ww,
                                                     cout), 'lpips.relu_pool^T')
                    else:
                        src = gfeat if i == 12 else g
                        check(lib.cagc_relu_pool_bwd(st, a.data_ptr(), None, src.data_ptr(), gz.data_ptr(), n, hh, ww, cout),
                              'lpips.relu^T')
                    del gfeat
                g = None
                p, _, tcd = _prep(ws[i], bs[i], 1.0, False, algo)
                out = _empty(n, hh, ww, cin, dev)
                # the activation below feeds this convolution only (no tap, no pool in between): its ReLU mask is applied
                # in the epilogue of the data-gradient convolution, which then hands the next layer its masked gradient
                plain_below = _MASK_EPILOGUE and (i - 1) not in TAPS and i not in POOL_BEFORE
                _vconv(st, gz, p.w_dgrad, None, out, n, hh, ww, cout, cin, False, tcd, f'lpips.conv{i}^T',
                       mask_ref=acts[i - 1] if plain_below else None)
                gz, g = (out, None) if plain_below else (None, out)
            a = acts[0]
            _, h, w, c0 = a.shape
            if gz is None:
                gz = _empty(n, h, w, c0, dev)
                check(lib.cagc_relu_pool_bwd(st, a.data_ptr(), None, g.data_ptr(), gz.data_ptr(), n, h, w, c0), 'lpips.relu^T')
            gimg = torch.empty((n, 3, h, w), device=dev, dtype=torch.float32)
            _timed('lpips_rgb_conv', 2.0 * n * h * w * 27 * c0, 4.0 * n * h * w * (3 + c0),
                   lambda: check(lib.cagc_rgb_conv3x3_bwd(st, gz.data_ptr(), ws[0].data_ptr(), mod._scale, gimg.data_ptr(),
                                                          n, h, w, c0, c0), 'lpips.conv1_1^T'))
        return gimg, None, None, None


def _vconv(st, x_buf, slab, bias_p, out, b, h, w, pin, pout, relu, tc, name, mask_ref=None):
    algo = config.ALGO_TCGEN05_TF32 if tc else config.ALGO_SIMT_FP32
    wsb, ws_bytes = conv_workspace(b, h, w, pout, out.device) if tc else (None, 0)
    if mask_ref is not None:       # data gradient with the ReLU mask of the activation it flows into (first b images)
        launch = lambda: check(lib.cagc_conv2d_mask_ws(st, x_buf.data_ptr(), slab.data_ptr(), mask_ref.data_ptr(),
                                                       out.data_ptr(), b, h, w, pin, pout, pout, 3, algo, ptr(wsb), ws_bytes),
                               name)
    else:
        launch = lambda: check(lib.cagc_conv2d_ws(st, x_buf.data_ptr(), slab.data_ptr(), ptr(bias_p), None, out.data_ptr(),
                                                  b, h, w, pin, pout, pout, 3, 0, int(relu), -1.0 if relu else 1.0, algo,
                                                  ptr(wsb), ws_bytes), name)
    _timed(f'lpips_conv[algo{algo}]', 2.0 * b * h * w * pin * pout * 9, 4.0 * b * h * w * (pin + pout), launch,
           shape=f'{pin}->{pout}x{h}x{w}')


class PerceptualLossVGG(nn.Module):
    """Drop-in for `lpips.PerceptualLoss(model='net-lin', net='vgg')` (lpips/__init__.py:13-41):
    `loss(pred, target, normalize=False) -> [N,1,1,1]`, frozen, CUDA only.

    conv_weights / conv_biases: the 13 VGG16 convolutions in `features` order; lin_weights: the five `lin{k}.model[1]`
    1x1 convolutions (lpips/weights/v0.1/vgg.pth), flattened to [C].
    """

    def __init__(self, conv_weights: Sequence[torch.Tensor], conv_biases: Sequence[torch.Tensor],
                 lin_weights: Sequence[torch.Tensor], shift=SHIFT, scale=SCALE):
        super().__init__()
        chans = [c for c in VGG16_CFG if c != 'M']
        if len(conv_weights) != 13 or len(conv_biases) != 13 or len(lin_weights) != 5:
            raise ValueError('PerceptualLossVGG: 13 convolutions and 5 lin layers expected')
        cin = 3
        for i, (wt, bt) in enumerate(zip(conv_weights, conv_biases)):
            if tuple(wt.shape) != (chans[i], cin, 3, 3) or tuple(bt.shape) != (chans[i],):
                raise ValueError(f'PerceptualLossVGG: conv {i} has shape {tuple(wt.shape)}, expected {(chans[i], cin, 3, 3)}')
            cin = chans[i]
        mk = lambda t: nn.Parameter(t.detach().clone().float().contiguous(), requires_grad=False)
        self.conv_weights = nn.ParameterList([mk(t) for t in conv_weights])
        self.conv_biases = nn.ParameterList([mk(t) for t in conv_biases])
        self.lin_weights = nn.ParameterList([mk(t.reshape(-1)) for t in lin_weights])
        for kk, i in enumerate(TAPS):
            if self.lin_weights[kk].numel() != chans[i]:
                raise ValueError(f'PerceptualLossVGG: lin{kk} has {self.lin_weights[kk].numel()} weights, expected {chans[i]}')
        self.shift, self.scale = tuple(float(v) for v in shift), tuple(float(v) for v in scale)

    # host-side float[3] arguments of the conv1_1 kernels (built per call: ctypes arrays do not pickle)
    @property
    def _shift(self):
        return _f3(self.shift)

    @property
    def _scale(self):
        return _f3(self.scale)

    @classmethod
    def from_reference(cls, percept_loss) -> 'PerceptualLossVGG':
        """Take the parameters of a reference `lpips.PerceptualLoss` / `DistModel` / `PNetLin` object (vgg, net-lin)."""
        net = percept_loss
        if not hasattr(net, 'scaling_layer') and not isinstance(getattr(net, 'model', ''), str):
            net = net.model                              # PerceptualLoss.model: DistModel
        if not hasattr(net, 'scaling_layer') and hasattr(net, 'net'):
            net = net.net                                # DistModel.net: PNetLin, possibly inside DataParallel
        if isinstance(net, nn.DataParallel):
            net = net.module
        if not hasattr(net, 'scaling_layer') or getattr(net, 'pnet_type', 'vgg') not in ('vgg', 'vgg16') \
                or not getattr(net, 'lpips', True) or getattr(net, 'spatial', False):
            raise ValueError('from_reference: a non-spatial net-lin VGG16 LPIPS network is required')
        convs = [m for s in (net.net.slice1, net.net.slice2, net.net.slice3, net.net.slice4, net.net.slice5)
                 for m in s if isinstance(m, nn.Conv2d)]
        lins = [lin.model[-1].weight for lin in (net.lin0, net.lin1, net.lin2, net.lin3, net.lin4)]
        out = cls([m.weight for m in convs], [m.bias for m in convs], lins,
                  shift=net.scaling_layer.shift.reshape(-1).tolist(), scale=net.scaling_layer.scale.reshape(-1).tolist())
        return out.to(convs[0].weight.device)

    def forward(self, pred, target, normalize=False):
        if normalize:                                    # lpips/__init__.py:36-38
            target = 2 * target - 1
            pred = 2 * pred - 1
        return _LpipsFn.apply(pred, target, self, config.conv_algo())
