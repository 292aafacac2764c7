"""Drop-in `model` module: StyleGAN2 Generator / Discriminator with the reference's public surface.

Same class names, constructor / forward signatures, attribute names and state_dict keys (and key
order) as the reference `model.py` (file:line cited per class), so that prune.py, train.py,
get_fid.py, get_ppl.py and the Util/ helpers bind to this implementation unmodified and reference
checkpoints load in both directions.  What differs is everything underneath: the modulated
convolution runs as fused sm_100a kernels with shared weights (style modulation on the operand load,
demodulation / noise / bias / leaky-ReLU in the epilogue), activations travel as channel-padded NHWC
buffers, ToRGB fuses bias and the upsampled skip connection -- see b200gan/modconv.py and DESIGN.md.

CUDA only: CPU tensors raise (the CPU restatement is oracle/stylegan2_oracle.py, test infrastructure).
"""
import math
import random

import torch
from torch import nn, autograd
from torch.nn import functional as F

from op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d
from b200gan import config as _cfg
from b200gan import modconv as _mc
from b200gan import dconv as _dc


import os as _os
_THROUGH = _os.environ.get('CAGC_TORGB_THROUGH', '1') != '0'   # tuning knob (development)


class PixelNorm(nn.Module):
    """x * rsqrt(mean(x^2, dim=1) + 1e-8)   (reference model.py:14-24)."""

    def forward(self, input):
        return input * torch.rsqrt(input.square().mean(dim=1, keepdim=True) + 1e-8)


def make_kernel(k):
    """Normalised 2-D FIR kernel from 1-D taps (outer product) or a 2-D array (reference model.py:27-35)."""
    k = torch.as_tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum()


def _resample_pads(n_taps, factor, kernel_size=None, up=True):
    """Padding arithmetic of reference model.py:46-51 (Upsample), :67-72 (Downsample),
    :207-221 (blur around a strided modulated conv)."""
    if kernel_size is None:
        p = n_taps - factor
        return ((p + 1) // 2 + factor - 1, p // 2) if up else ((p + 1) // 2, p // 2)
    if up:
        p = (n_taps - factor) - (kernel_size - 1)
        return (p + 1) // 2 + factor - 1, p // 2 + 1
    p = (n_taps - factor) + (kernel_size - 1)
    return (p + 1) // 2, p // 2


class Upsample(nn.Module):
    """Zero-insertion x`factor` + FIR (reference model.py:38-56)."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        fir = make_kernel(kernel) * (factor ** 2)
        self.register_buffer('kernel', fir)
        self.pad = _resample_pads(fir.shape[0], factor, up=True)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):
    """FIR + decimation by `factor` (reference model.py:59-77)."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        fir = make_kernel(kernel)
        self.register_buffer('kernel', fir)
        self.pad = _resample_pads(fir.shape[0], factor, up=False)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


class Blur(nn.Module):
    """FIR low-pass with explicit pads (reference model.py:80-96)."""

    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        fir = make_kernel(kernel)
        if upsample_factor > 1:
            fir = fir * (upsample_factor ** 2)
        self.register_buffer('kernel', fir)
        self.pad = pad

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


class EqualConv2d(nn.Module):
    """Equalised-lr convolution of the discriminator (reference model.py:99-134); plain library conv."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride = stride
        self.padding = padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    def forward(self, input):
        w = _mc.cached_frozen(self.weight, ('eqconv', self.scale), lambda: self.weight * self.scale)
        return F.conv2d(input, w, bias=self.bias, stride=self.stride, padding=self.padding)

    def __repr__(self):
        o, i, k, _ = self.weight.shape
        return f'{self.__class__.__name__}({i}, {o}, {k}, stride={self.stride}, padding={self.padding})'


class EqualLinear(nn.Module):
    """Equalised-lr linear layer, optional fused leaky ReLU (reference model.py:137-171).
    The GEMM is a plain library call (M = batch); the bias + activation is our fused kernel."""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    # Generator marks its own layers: library GEMM + ONE epilogue kernel (first-order autograd only).  Everything
    # else (the discriminator head, any second-order path) runs the differentiable torch composition.
    _fused = False

    def forward(self, input):
        if self._fused and not _cfg.is_second_order() and input.is_cuda:
            return _mc.equal_linear(input, self.weight, self.bias, self.scale, self.lr_mul, bool(self.activation))
        w = _mc.cached_frozen(self.weight, ('eqlin', self.scale), lambda: self.weight * self.scale)
        if self.activation:
            return fused_leaky_relu(F.linear(input, w), self.bias * self.lr_mul)
        return F.linear(input, w, bias=self.bias * self.lr_mul)

    def __repr__(self):
        return f'{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})'


class ScaledLeakyReLU(nn.Module):
    """leaky_relu(x) * sqrt(2) without bias (reference model.py:174-183)."""

    def __init__(self, negative_slope=0.2):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, input):
        return fused_leaky_relu(input, None, self.negative_slope, math.sqrt(2))


class ModulatedConv2d(nn.Module):
    """Style-modulated, demodulated convolution (reference model.py:186-289).

    Parameters / attributes are those of the reference (`weight [1,O,I,k,k]`, `modulation`,
    `scale`, `demodulate`, `blur`, ...).  forward() never materialises per-sample weights: it
    computes s = modulation(style), the demodulation coefficients d[b,o], and hands both to the
    fused kernels together with the shared weight.
    """

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        if upsample:
            self.blur = Blur(blur_kernel, pad=_resample_pads(len(blur_kernel), 2, kernel_size, up=True),
                             upsample_factor=2)
        if downsample:
            self.blur = Blur(blur_kernel, pad=_resample_pads(len(blur_kernel), 2, kernel_size, up=False))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate

    def __repr__(self):
        return (f'{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, '
                f'upsample={self.upsample}, downsample={self.downsample})')

    def _run(self, input, style, noise=None, noise_weight=None, act_bias=None, act=False, s_p=None):
        """Shared body of ModulatedConv2d.forward and the fused StyledConv.forward.  `s_p`: style scalars
        already computed (zero padded to the channel pitch) by the generator's one-launch modulation."""
        if self.downsample or _cfg.is_second_order():
            s = self.modulation(style)                                            # [B, I]
            d = _mc.demod_coefficients(s, self.weight, self.scale, self.eps) if self.demodulate else None
            if self.downsample:
                # not used by the generator or the discriminator of this repo; differentiable composite
                w = (self.weight[0] * self.scale)
                x = self.blur(input * s[:, :, None, None])
                out = F.conv2d(x, w, stride=2)
                if d is not None:
                    out = out * d[:, :, None, None]
                if noise is not None:
                    out = out + noise_weight * noise
                if act:
                    out = fused_leaky_relu(out, act_bias)
                return out, s
            fir = self.blur.kernel if self.upsample else None
            pad = self.blur.pad if self.upsample else (0, 0)
            out = _mc.styled_conv_composite(input, s, d, self.weight, noise, noise_weight, act_bias, self.scale,
                                            upsample=self.upsample, fir=fir, pad=pad, act=act)
            return out, s
        if s_p is None:
            s_p = _mc.style_affine(style.unsqueeze(1), [self.modulation], [0])[0]
        fir = self.blur.kernel if self.upsample else None
        pad = self.blur.pad if self.upsample else (0, 0)
        out = _mc.styled_conv(input, s_p, self.weight, noise, noise_weight, act_bias, self.scale,
                              demodulate=self.demodulate, eps=self.eps, upsample=self.upsample, fir=fir, pad=pad,
                              act=act)
        return out, s_p[:, :self.in_channel]

    def forward(self, input, style, return_style_scalars=False):
        out, s = self._run(input, style)
        if return_style_scalars:
            return out, s.reshape(s.shape[0], 1, self.in_channel, 1, 1)
        return out


class NoiseInjection(nn.Module):
    """image + weight * noise, fresh N(0,1) noise when none is given (reference model.py:292-303)."""

    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            batch, _, height, width = image.shape
            noise = image.new_empty(batch, 1, height, width).normal_()
        return image + self.weight * noise


class ConstantInput(nn.Module):
    """Learned 4x4 constant repeated over the batch (reference model.py:306-320)."""

    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class StyledConv(nn.Module):
    """ModulatedConv2d -> NoiseInjection -> FusedLeakyReLU (reference model.py:323-367), executed
    as one fused kernel (two for the upsampling variant)."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, return_style_scalars=False, noise=None, _s=None):
        if noise is None:
            # same RNG consumption as the reference: one normal_() per layer, in call order (model.py:299-301)
            b, _, h, w = input.shape
            f = 2 if self.conv.upsample else 1
            noise = input.new_empty(b, 1, h * f, w * f).normal_()
        act = self.activate
        if act.negative_slope != 0.2 or abs(act.scale - math.sqrt(2)) > 1e-12 or \
                (noise.requires_grad and torch.is_grad_enabled()):
            # non-default activation constants, or a caller that optimises the noise maps themselves (the projector,
            # Miscellaneous/Image2StyleGAN_util.py:54, sets noise.requires_grad): unfused, differentiable tail
            out, s = self.conv._run(input, style, s_p=_s)
            out = act(self.noise(out, noise=noise))
        else:
            out, s = self.conv._run(input, style, noise=noise, noise_weight=self.noise.weight, act_bias=act.bias,
                                    act=True, s_p=_s)
        if return_style_scalars:
            return out, s.reshape(s.shape[0], 1, self.conv.in_channel, 1, 1)
        return out


class ToRGB(nn.Module):
    """1x1 modulated conv (no demodulation) + bias + upsampled skip (reference model.py:370-395)."""

    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, input, style, skip=None, return_style_scalars=False, _s=None, _through=False):
        """_through (private, first-order fused path only): also return the input activation routed through this
        node, so that the gradients of its two consumers are summed inside the ToRGB backward kernel."""
        conv = self.conv
        if skip is not None:
            up = self.upsample
            if up.factor != 2:
                raise RuntimeError('ToRGB: only factor-2 skip upsampling is implemented')
            fir, pad = up.kernel, up.pad
        else:
            fir, pad = None, (0, 0)
        if _cfg.is_second_order():
            s = conv.modulation(style)
            out = _mc.to_rgb_composite(input, s, conv.weight, self.bias, skip, conv.scale, fir=fir, pad=pad)
        else:
            s_p = _s if _s is not None else _mc.style_affine(style.unsqueeze(1), [conv.modulation], [0])[0]
            out = _mc.to_rgb(input, s_p, conv.weight, self.bias, skip, conv.scale, fir=fir, pad=pad,
                             passthrough=_through)
            s = s_p[:, :conv.in_channel]
            if _through:
                rgb, x_through = out
                if return_style_scalars:
                    return rgb, s.reshape(s.shape[0], 1, conv.in_channel, 1, 1), x_through
                return rgb, x_through
        if return_style_scalars:
            return out, s.reshape(s.shape[0], 1, conv.in_channel, 1, 1)
        return out


class Generator(nn.Module):
    """StyleGAN2 generator, optionally channel-pruned through `generator_net_shape`
    (reference model.py:398-666)."""

    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01,
                 generator_net_shape=None):
        super().__init__()
        self.size = size
        self.style_dim = style_dim

        mapping = [PixelNorm()]
        for _ in range(n_mlp):
            mapping.append(EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation='fused_lrelu'))
        self.style = nn.Sequential(*mapping)
        for m in mapping[1:]:
            m._fused = True

        self.channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier,
                         128: 128 * channel_multiplier, 256: 64 * channel_multiplier,
                         512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.n_latent = self.log_size * 2 - 2

        # widths[i] = input channels of styled conv i; widths[-1] = output channels of the last one
        if generator_net_shape is None:
            widths = [self.channels[4], self.channels[4]]
            for r in range(3, self.log_size + 1):
                widths += [self.channels[2 ** r]] * 2
        else:
            widths = list(generator_net_shape)

        self.input = ConstantInput(widths[0])
        self.conv1 = StyledConv(widths[0], widths[1], 3, style_dim, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(widths[1], style_dim, upsample=False)

        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer(f'noise_{layer_idx}', torch.randn(1, 1, 2 ** res, 2 ** res))

        for blk in range(1, len(widths) // 2):
            c_in, c_mid, c_out = widths[2 * blk - 1], widths[2 * blk], widths[2 * blk + 1]
            self.convs.append(StyledConv(c_in, c_mid, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(StyledConv(c_mid, c_out, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(c_out, style_dim))

    def make_noise(self):
        device = self.input.input.device
        noises = [torch.randn(1, 1, 4, 4, device=device)]
        for r in range(3, self.log_size + 1):
            for _ in range(2):
                noises.append(torch.randn(1, 1, 2 ** r, 2 ** r, device=device))
        return noises

    def mean_latent(self, n_latent):
        z = torch.randn(n_latent, self.style_dim, device=self.input.input.device)
        return self.style(z).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    def forward(self, noise_z, return_latents=False, inject_index=None, truncation=1, truncation_latent=None,
                latent_styles=None, input_is_latent=False, noise=None, randomize_noise=True, PPL_regularize=False,
                return_rgb_list=False, return_style_scalars=False):
        if PPL_regularize and not _cfg.is_second_order():
            with _cfg.second_order():
                return self.forward(noise_z, return_latents, inject_index, truncation, truncation_latent,
                                    latent_styles, input_is_latent, noise, randomize_noise, True,
                                    return_rgb_list, return_style_scalars)

        if input_is_latent:
            styles = latent_styles
        elif len(noise_z) == 2 and noise_z[0].shape == noise_z[1].shape and noise_z[0].ndim == 2:
            # style mixing: both latents through the mapping network in one pass
            styles = list(self.style(torch.cat([noise_z[0], noise_z[1]], 0)).chunk(2, 0))
        else:
            styles = [self.style(z) for z in noise_z]

        if noise is None:
            if randomize_noise:
                noise = [None] * self.num_layers
            else:
                noise = [getattr(self.noises, f'noise_{i}') for i in range(self.num_layers)]

        if truncation < 1:
            styles = [truncation_latent + truncation * (w - truncation_latent) for w in styles]

        if len(styles) < 2:
            inject_index = self.n_latent
            latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1) if styles[0].ndim < 3 else styles[0]
        else:
            if inject_index is None:
                inject_index = random.randint(1, self.n_latent - 1)
            latent = torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                                styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)], 1)

        scalars = []

        # style modulation of every layer in ONE launch (the reference runs one EqualLinear per layer,
        # model.py:248): layer order = execution order, latent rows as in model.py:618-644
        if _cfg.is_second_order():
            s_of = {}
        else:
            mods, rows = [self.conv1.conv, self.to_rgb1.conv], [0, 1]
            for blk, to_rgb in enumerate(self.to_rgbs):
                mods += [self.convs[2 * blk].conv, self.convs[2 * blk + 1].conv, to_rgb.conv]
                rows += [1 + 2 * blk, 2 + 2 * blk, 3 + 2 * blk]
            s_all = _mc.style_affine(latent, [m.modulation for m in mods], rows)
            s_of = {id(m): sp for m, sp in zip(mods, s_all)}

        def styled(layer, x, w_lat, nz):
            if return_style_scalars:
                y, sc = layer(x, w_lat, True, noise=nz, _s=s_of.get(id(layer.conv)))
                scalars.append(sc)
                return y
            return layer(x, w_lat, noise=nz, _s=s_of.get(id(layer.conv)))

        # first-order fused path: an activation that feeds both a ToRGB and the next block travels THROUGH the ToRGB
        # node, which sums the two gradients in its backward kernel (no separate accumulation pass)
        through = not _cfg.is_second_order() and torch.is_grad_enabled() and _THROUGH
        out = self.input(latent)
        out = styled(self.conv1, out, latent[:, 0], noise[0])
        if through and len(self.to_rgbs) > 0:
            skip, out = self.to_rgb1(out, latent[:, 1], _s=s_of.get(id(self.to_rgb1.conv)), _through=True)
        else:
            skip = self.to_rgb1(out, latent[:, 1], _s=s_of.get(id(self.to_rgb1.conv)))
        rgbs = [skip]

        i = 1
        n_blocks = len(self.to_rgbs)
        for blk, to_rgb in enumerate(self.to_rgbs):
            out = styled(self.convs[2 * blk], out, latent[:, i], noise[1 + 2 * blk])
            out = styled(self.convs[2 * blk + 1], out, latent[:, i + 1], noise[2 + 2 * blk])
            thr = through and blk + 1 < n_blocks
            if return_style_scalars and (i + 3) == latent.shape[1]:   # only the last ToRGB reports (model.py:636-638)
                res = to_rgb(out, latent[:, i + 2], skip, True, _s=s_of.get(id(to_rgb.conv)), _through=thr)
                skip, sc = res[0], res[1]
                scalars.append(sc)
                if thr:
                    out = res[2]
            elif thr:
                skip, out = to_rgb(out, latent[:, i + 2], skip, _s=s_of.get(id(to_rgb.conv)), _through=True)
            else:
                skip = to_rgb(out, latent[:, i + 2], skip, _s=s_of.get(id(to_rgb.conv)))
            rgbs.append(skip)
            i += 2

        image = skip
        if not PPL_regularize:
            ret = rgbs if return_rgb_list else image
            return (ret, scalars) if return_style_scalars else ret

        # path-length regulariser quantity (reference model.py:661-666)
        pl_noise = torch.randn_like(image) / math.sqrt(image.shape[2] * image.shape[3])
        grad, = autograd.grad(outputs=(image * pl_noise).sum(), inputs=latent, create_graph=True)
        path_lengths = torch.sqrt(grad.pow(2).sum(2).mean(1))
        return image, path_lengths


# --------------------------------------------------------------------------------------------
# Discriminator (reference model.py:670-798).  Frozen (the KD generator step): every block runs on the
# package's own convolution engines with fused epilogues (b200gan/dconv.py).  Trainable (D steps, R1
# double backward -- outside the benchmarked path): differentiable composition of library convolutions
# with our upfirdn2d (Blur) and fused bias+leaky-ReLU kernels.
# --------------------------------------------------------------------------------------------
class ConvLayer(nn.Sequential):
    def __init__(self, in_channel, out_channel, kernel_size, downsample=False, blur_kernel=[1, 3, 3, 1], bias=True,
                 activate=True):
        layers = []
        if downsample:
            layers.append(Blur(blur_kernel, pad=_resample_pads(len(blur_kernel), 2, kernel_size, up=False)))
            stride, self.padding = 2, 0
        else:
            stride, self.padding = 1, kernel_size // 2
        layers.append(EqualConv2d(in_channel, out_channel, kernel_size, padding=self.padding, stride=stride,
                                  bias=bias and not activate))
        if activate:
            layers.append(FusedLeakyReLU(out_channel) if bias else ScaledLeakyReLU(0.2))
        super().__init__(*layers)


class ResBlock(nn.Module):
    def __init__(self, in_channel, out_channel, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.conv1 = ConvLayer(in_channel, in_channel, 3)
        self.conv2 = ConvLayer(in_channel, out_channel, 3, downsample=True)
        self.skip = ConvLayer(in_channel, out_channel, 1, downsample=True, activate=False, bias=False)

    def forward(self, input):
        if input.is_cuda and not _cfg.is_second_order() and _dc.res_block_eligible(self):
            # frozen discriminator (the generator step, train.py:286-287): the whole block is one autograd node on the
            # package's convolution engines, bias / activation / residual merge in the epilogues (b200gan/dconv.py)
            return _dc.res_block(input, self)
        out = self.conv1(input)
        c2, sk = self.conv2, self.skip
        if isinstance(c2[-1], FusedLeakyReLU) and len(c2) == 3 and len(sk) == 2 and sk[1].bias is None:
            # (conv2(x) + skip(x)) / sqrt2 with the 1/sqrt2 folded into the activation gain of conv2 and into
            # the (linear, bias-free) skip convolution's weight scale: one add instead of add + scale, forward
            # and backward (reference model.py:731-737 computes the same quantity)
            r = 1 / math.sqrt(2)
            act = c2[2]
            out = fused_leaky_relu(c2[1](c2[0](out)), act.bias, act.negative_slope, act.scale * r)
            conv = sk[1]
            w = _mc.cached_frozen(conv.weight, ('eqconv', conv.scale * r), lambda: conv.weight * (conv.scale * r))
            return out + F.conv2d(sk[0](input), w, stride=conv.stride, padding=conv.padding)
        return (self.conv2(out) + self.skip(input)) / math.sqrt(2)


class Discriminator(nn.Module):
    def __init__(self, size, channel_multiplier=2, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier,
                    128: 128 * channel_multiplier, 256: 64 * channel_multiplier,
                    512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}
        log_size = int(math.log(size, 2))
        convs = [ConvLayer(3, channels[size], 1)]
        in_channel = channels[size]
        for r in range(log_size, 2, -1):
            out_channel = channels[2 ** (r - 1)]
            convs.append(ResBlock(in_channel, out_channel, blur_kernel))
            in_channel = out_channel
        self.convs = nn.Sequential(*convs)
        self.stddev_group = 4
        self.stddev_feat = 1
        self.final_conv = ConvLayer(in_channel + 1, channels[4], 3)
        self.final_linear = nn.Sequential(
            EqualLinear(channels[4] * 4 * 4, channels[4], activation='fused_lrelu'),
            EqualLinear(channels[4], 1),
        )

    def _own_engines(self, input):
        """True when this forward can run on the package's convolution engines: CUDA input, first-order autograd,
        every parameter frozen (train.py:286-287) and the standard layer layout."""
        if not input.is_cuda or _cfg.is_second_order() or input.shape[1] > 4:
            return False
        first, last = self.convs[0], self.final_conv
        if not (len(first) == 2 and isinstance(first[0], EqualConv2d) and isinstance(first[1], FusedLeakyReLU)
                and first[0].weight.shape[-1] == 1 and len(last) == 2 and isinstance(last[1], FusedLeakyReLU)):
            return False
        return all(not p.requires_grad for p in self.parameters())

    def forward(self, input):
        own = self._own_engines(input)
        if own:
            out = _dc.from_rgb(input, self.convs[0][0], self.convs[0][1])
            for blk in list(self.convs)[1:]:
                out = blk(out)
        else:
            out = self.convs(input)
        batch, channel, height, width = out.shape
        group = min(batch, self.stddev_group)
        # minibatch standard deviation feature (reference model.py:783-791)
        sd = out.reshape(group, -1, self.stddev_feat, channel // self.stddev_feat, height, width)
        sd = torch.sqrt(sd.var(0, unbiased=False) + 1e-8)
        sd = sd.mean([2, 3, 4], keepdim=True).squeeze(2)
        sd = sd.repeat(group, 1, height, width)
        out = torch.cat([out, sd], 1)
        if own:
            out = _dc.conv_act(out, self.final_conv[0], self.final_conv[1])
            out = out.reshape(batch, -1)
            for lin in self.final_linear:
                out = _mc.equal_linear(out, lin.weight, lin.bias, lin.scale, lin.lr_mul, bool(lin.activation))
            return out
        out = self.final_conv(out)
        out = out.reshape(batch, -1)
        return self.final_linear(out)
