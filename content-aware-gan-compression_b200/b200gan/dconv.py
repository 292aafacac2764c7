"""Discriminator blocks (reference model.py:670-798) on the package's own convolution engines.

The reference runs every ConvLayer as [Blur ->] F.conv2d (cuDNN) -> FusedLeakyReLU and the ResBlock tail as
(conv2(out) + skip(x)) / sqrt2.  Here a frozen discriminator (the KD generator step: train.py:286-287,
requires_grad(D, False)) runs as ONE autograd node per block, on NHWC-p buffers:

    forward                                                   backward (data gradient only: D is frozen)
    a1 = lrelu(conv3x3(x, W1) + b1) * sqrt2       [1 launch]  gs  = conv1x1^T(g, Ws / sqrt2)
    t  = Blur(a1, pad (2,2))                      [1]         gxs = FIR-up2(gs)            (adjoint of FIR-down2)
    a2 = lrelu(conv3x3 stride 2 (t, W2) + b2)     [1]         gz2 = g * mask(a2);  gt = convT stride 2 (gz2, W2)
    s  = FIR-down2(x)   (Blur pad (1,1) at even positions) [1] gz1 = Blur^T(gt) * sqrt2 * mask(a1)
    y  = conv1x1(s, Ws / sqrt2) + a2              [1]         gx  = conv3x3^T(gz1, W1) + gxs   (sum in the epilogue)

so bias, activation, the 1/sqrt2 of the residual merge and both gradient sums live in convolution epilogues and
no elementwise pass (fused_bias_act, at::add, at::mul) remains.  Engines: `config.conv_algo()` -- tcgen05 TF32
(TMA-fed implicit GEMM, conv_tc.cu) or the exact-fp32 SIMT engine.

A discriminator whose parameters require gradients (train.py's D_Loss_BackProp / R1 steps, outside the
benchmarked path) keeps the differentiable composition of library convolutions + our upfirdn2d / fused_leaky_relu
in model.py.
"""
from __future__ import annotations

import math

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ._lib import lib, check, stream_of, require_cuda, ptr, host_taps, fir_nhwc, conv_workspace
from . import config
from .modconv import (_frozen, _weight_prep, _use_tc, _flipped, as_nhwc_buf, nhwc_view, pitch_of, _timed)

SQRT2 = math.sqrt(2.0)


def frozen(*params) -> bool:
    return all(p is None or not (p.requires_grad or p.grad_fn is not None) for p in params)


def _prep(weight: torch.Tensor, bias, wscale: float, strided: bool, algo: int):
    """Operand slabs of a frozen EqualConv2d weight [O,I,k,k] (cached per parameter identity / version):
    .w_fwd (forward), .w_dgrad (data gradient: flipped taps for the same-size conv, plain for the stride-2 conv whose
    gradient is a transposed convolution), .bias_p (zero padded to the channel pitch)."""
    cout, cin, k, _ = weight.shape
    pin, pout = pitch_of(cin), pitch_of(cout)
    tc_f, tc_d = _use_tc(algo, pin), _use_tc(algo, pout)
    tag = ('dconv', wscale, strided, tc_f, tc_d, None if bias is None else (id(bias), bias._version))

    def make():
        w5 = weight.detach().reshape(1, cout, cin, k, k)
        return _weight_prep(w5, bias, wscale, strided, tc_f, tc_d, pin, pout, False), tc_f, tc_d
    return _frozen.get(weight, tag, make)


def _conv(st, x_buf, slab, bias_p, residual, out, b, h, w, pin, pout, cout, k, mode, act, gain, tc, name):
    algo = config.ALGO_TCGEN05_TF32 if tc else config.ALGO_SIMT_FP32
    ho, wo = (h, w) if mode == 0 else ((h - k) // 2 + 1, (w - k) // 2 + 1)
    flops = 2.0 * b * ho * wo * pin * cout * k * k
    ws, ws_bytes = conv_workspace(b, ho, wo, pout, out.device) if tc else (None, 0)
    _timed(f'dconv[algo{algo}]', flops, 4.0 * b * (h * w * pin + ho * wo * pout),
           lambda: check(lib.cagc_conv2d_ws(st, x_buf.data_ptr(), slab.data_ptr(), ptr(bias_p), ptr(residual),
                                            out.data_ptr(), b, h, w, pin, pout, cout, k, mode, int(act), gain, algo,
                                            ptr(ws), ws_bytes), name),
           shape=f'{pin}->{cout}x{ho}x{wo}k{k}s{1 + mode}')


def _empty(b, h, w, p, dev):
    return torch.empty((b, h, w, p), device=dev, dtype=torch.float32)


def _taps(fir: torch.Tensor):
    t = host_taps(fir)
    if t is None:
        raise RuntimeError('FIR taps of a discriminator Blur must be seen once outside CUDA-graph capture '
                           '(KDStep.capture warms up eagerly)')
    return t


class _ResBlockFn(Function):
    """y = (conv2(conv1(x)) + skip(x)) / sqrt2 of reference model.py:719-737, frozen parameters."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, ws, fir2, pad2, firs, pads, scale1, scale2, scales, algo):
        require_cuda(x, 'ResBlock')
        b, cin, h, w = x.shape
        cout = w2.shape[0]
        pin, pout = pitch_of(cin), pitch_of(cout)
        dev = x.device
        r = 1.0 / SQRT2
        with torch.cuda.device(dev):
            st = stream_of(x)
            xb = as_nhwc_buf(x.detach())
            p1, tc1f, tc1d = _prep(w1, b1, scale1, False, algo)
            p2, tc2f, tc2d = _prep(w2, b2, scale2, True, algo)
            p3, tc3f, tc3d = _prep(ws, None, scales * r, False, algo)
            a1 = _empty(b, h, w, pin, dev)
            _conv(st, xb, p1.w_fwd, p1.bias_p, None, a1, b, h, w, pin, pin, cin, 3, 0, True, SQRT2, tc1f, 'resblock.conv1')
            # Blur before the stride-2 3x3 conv: pad (2, 2) (model.py:683-689) -> (h+1) x (w+1)
            ht, wt = h + pad2[0] + pad2[1] - 3, w + pad2[0] + pad2[1] - 3
            t = _empty(b, ht, wt, pin, dev)
            _timed('fir_nhwc', 0.0, 4.0 * b * cin * (h * w + ht * wt),
                   lambda: fir_nhwc(st, a1.data_ptr(), fir2, None, None, None, None, t.data_ptr(), b, h, w, pin, cin,
                                    (pad2[0], pad2[1], pad2[0], pad2[1]), 0, 0, 'resblock.blur'))
            ho, wo = (ht - 3) // 2 + 1, (wt - 3) // 2 + 1
            a2 = _empty(b, ho, wo, pout, dev)
            # activation gain sqrt2 * (1/sqrt2) = 1: the residual merge's 1/sqrt2 is folded in (and into Ws)
            _conv(st, t, p2.w_fwd, p2.bias_p, None, a2, b, ht, wt, pin, pout, cout, 3, 1, True, 1.0, tc2f, 'resblock.conv2')
            del t
            # skip: Blur(pad (1,1)) evaluated at the even positions only, then the 1x1 conv adds a2 in its epilogue
            s = _empty(b, ho, wo, pin, dev)
            check(lib.cagc_fir_resample_nhwc(st, xb.data_ptr(), _taps(firs), s.data_ptr(), b, h, w, pin, 1, 2,
                                             pads[0], pads[1]), 'resblock.fir_down2')
            y = _empty(b, ho, wo, pout, dev)
            _conv(st, s, p3.w_fwd, None, a2, y, b, ho, wo, pin, pout, cout, 1, 0, False, 1.0, tc3f, 'resblock.skip')
        ctx.save_for_backward(a1, a2, w1, b1, w2, b2, ws, fir2, firs)
        ctx.cfg = (b, cin, cout, h, w, ht, wt, ho, wo, pin, pout, pad2, pads, scale1, scale2, scales, algo)
        return nhwc_view(y, cout)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        a1, a2, w1, b1, w2, b2, ws, fir2, firs = ctx.saved_tensors
        (b, cin, cout, h, w, ht, wt, ho, wo, pin, pout, pad2, pads, scale1, scale2, scales, algo) = ctx.cfg
        if not ctx.needs_input_grad[0]:
            return (None,) * 14
        dev = a1.device
        r = 1.0 / SQRT2
        with torch.cuda.device(dev):
            st = stream_of(a1)
            gb = as_nhwc_buf(g)
            p1, tc1f, tc1d = _prep(w1, b1, scale1, False, algo)
            p2, tc2f, tc2d = _prep(w2, b2, scale2, True, algo)
            p3, tc3f, tc3d = _prep(ws, None, scales * r, False, algo)
            # skip branch: 1x1 conv transposed, then the adjoint of the decimating FIR
            gs = _empty(b, ho, wo, pin, dev)
            _conv(st, gb, p3.w_dgrad, None, None, gs, b, ho, wo, pout, pin, pin, 1, 0, False, 1.0, tc3d, 'resblock.skip^T')
            gxs = _empty(b, h, w, pin, dev)
            # gradient pads of upfirdn2d (op/upfirdn2d.py:111-116) for up = 1, down = 2, 4 taps
            gp0 = 4 - pads[0] - 1
            gp1 = h - ho * 2 + pads[0] - 1 + 1
            check(lib.cagc_fir_resample_nhwc(st, gs.data_ptr(), _taps(_flipped(firs)), gxs.data_ptr(), b, ho, wo, pin,
                                             2, 1, gp0, gp1), 'resblock.fir_up2')
            del gs
            # main branch
            gz2 = torch.empty_like(a2)
            check(lib.cagc_act_mask_nhwc(st, gb.data_ptr(), a2.data_ptr(), gz2.data_ptr(), gz2.numel(), 1.0),
                  'resblock.act2^T')
            gt = _empty(b, ht, wt, pin, dev)
            algo2 = config.ALGO_TCGEN05_TF32 if tc2d else config.ALGO_SIMT_FP32
            ws, ws_bytes = conv_workspace(b, ht, wt, pin, dev) if tc2d else (None, 0)
            _timed(f'dconv_up[algo{algo2}]', 2.0 * b * ho * wo * pin * cout * 9, 4.0 * b * (ho * wo * pout + ht * wt * pin),
                   lambda: check(lib.cagc_conv_up_ws(st, gz2.data_ptr(), p2.w_dgrad.data_ptr(), None, gt.data_ptr(), b, ho, wo,
                                                     pout, pin, 3, algo2, ptr(ws), ws_bytes), 'resblock.conv2^T'))
            del ws
            del gz2
            q0, q1 = 4 - pad2[0] - 1, h - ht + pad2[0]
            firf = _flipped(fir2)
            gz1 = _empty(b, h, w, pin, dev)
            # Blur^T and the backward of conv1's activation in ONE pass (mask from the sign of a1, gain sqrt2)
            rc = _timed('fir_nhwc', 0.0, 4.0 * b * cin * (2 * h * w + ht * wt),
                        lambda: lib.cagc_fir_nhwc_mask(st, gt.data_ptr(), firf.data_ptr(), _taps(firf), a1.data_ptr(), SQRT2,
                                                       gz1.data_ptr(), b, ht, wt, pin, pin, 4, 4, q0, q1, q0, q1))
            if rc != 0:        # shape outside the row-ring kernel: two passes
                ga1 = _empty(b, h, w, pin, dev)
                _timed('fir_nhwc', 0.0, 4.0 * b * cin * (h * w + ht * wt),
                       lambda: fir_nhwc(st, gt.data_ptr(), firf, None, None, None, None, ga1.data_ptr(), b, ht, wt,
                                        pin, pin, (q0, q1, q0, q1), 0, 0, 'resblock.blur^T'))
                check(lib.cagc_act_mask_nhwc(st, ga1.data_ptr(), a1.data_ptr(), gz1.data_ptr(), gz1.numel(), SQRT2),
                      'resblock.act1^T')
                del ga1
            del gt
            gx = _empty(b, h, w, pin, dev)
            _conv(st, gz1, p1.w_dgrad, None, gxs, gx, b, h, w, pin, pin, cin, 3, 0, False, 1.0, tc1d, 'resblock.conv1^T')
        return (nhwc_view(gx, cin),) + (None,) * 13


def res_block(x, block, algo=None):
    """block: model.ResBlock with frozen parameters."""
    c1, c2, sk = block.conv1, block.conv2, block.skip
    conv1, act1 = c1[0], c1[1]
    blur2, conv2, act2 = c2[0], c2[1], c2[2]
    blurs, convs = sk[0], sk[1]
    return _ResBlockFn.apply(x, conv1.weight, act1.bias, conv2.weight, act2.bias, convs.weight,
                             blur2.kernel, tuple(blur2.pad), blurs.kernel, tuple(blurs.pad),
                             conv1.scale, conv2.scale, convs.scale, config.conv_algo() if algo is None else algo)


def res_block_eligible(block) -> bool:
    from model import FusedLeakyReLU, Blur, EqualConv2d
    c1, c2, sk = block.conv1, block.conv2, block.skip
    if not (len(c1) == 2 and len(c2) == 3 and len(sk) == 2):
        return False
    ok = isinstance(c1[0], EqualConv2d) and isinstance(c1[1], FusedLeakyReLU) and isinstance(c2[0], Blur) and \
        isinstance(c2[1], EqualConv2d) and isinstance(c2[2], FusedLeakyReLU) and isinstance(sk[0], Blur) and \
        isinstance(sk[1], EqualConv2d)
    if not ok:
        return False
    for a in (c1[1], c2[2]):
        if a.negative_slope != 0.2 or abs(a.scale - SQRT2) > 1e-12:
            return False
    if c1[0].weight.shape[-1] != 3 or c2[1].weight.shape[-1] != 3 or sk[1].weight.shape[-1] != 1:
        return False
    if c1[0].stride != 1 or c2[1].stride != 2 or sk[1].stride != 2 or sk[1].bias is not None:
        return False
    if tuple(c2[0].kernel.shape) != (4, 4) or tuple(sk[0].kernel.shape) != (4, 4):
        return False
    if c1[0].weight.shape[1] % 8 or c2[1].weight.shape[0] % 8:
        return False
    return frozen(c1[0].weight, c1[0].bias, c1[1].bias, c2[1].weight, c2[1].bias, c2[2].bias, sk[1].weight)


class _ConvActFn(Function):
    """y = [lrelu * sqrt2](conv_kxk(x, W * scale, same padding) + bias), frozen parameters (final_conv, model.py:793)."""

    @staticmethod
    def forward(ctx, x, weight, bias, scale, act, algo):
        require_cuda(x, 'ConvLayer')
        b, cin, h, w = x.shape
        cout, _, k, _ = weight.shape
        pin, pout = pitch_of(cin), pitch_of(cout)
        dev = x.device
        with torch.cuda.device(dev):
            st = stream_of(x)
            xb = as_nhwc_buf(x.detach())
            p, tcf, tcd = _prep(weight, bias, scale, False, algo)
            y = _empty(b, h, w, pout, dev)
            _conv(st, xb, p.w_fwd, p.bias_p, None, y, b, h, w, pin, pout, cout, k, 0, act, SQRT2, tcf, 'convlayer')
        ctx.save_for_backward(y if act else None, weight, bias)
        ctx.cfg = (b, cin, cout, h, w, k, pin, pout, scale, act, algo)
        return nhwc_view(y, cout)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        y, weight, bias = ctx.saved_tensors
        b, cin, cout, h, w, k, pin, pout, scale, act, algo = ctx.cfg
        if not ctx.needs_input_grad[0]:
            return (None,) * 6
        dev = g.device
        with torch.cuda.device(dev):
            st = stream_of(g)
            gb = as_nhwc_buf(g)
            p, tcf, tcd = _prep(weight, bias, scale, False, algo)
            if act:
                gz = torch.empty_like(y)
                check(lib.cagc_act_mask_nhwc(st, gb.data_ptr(), y.data_ptr(), gz.data_ptr(), gz.numel(), SQRT2), 'convlayer.act^T')
            else:
                gz = gb
            gx = _empty(b, h, w, pin, dev)
            _conv(st, gz, p.w_dgrad, None, None, gx, b, h, w, pout, pin, cin, k, 0, False, 1.0, tcd, 'convlayer^T')
        return (nhwc_view(gx, cin),) + (None,) * 5


def conv_act(x, conv, act_module, algo=None):
    return _ConvActFn.apply(x, conv.weight, act_module.bias if act_module is not None else conv.bias, conv.scale,
                            act_module is not None, config.conv_algo() if algo is None else algo)


class _FromRGBFn(Function):
    """ConvLayer(3, C, 1) (model.py:756): 1x1 conv over the image channels + bias + lrelu * sqrt2, image in any
    strided layout, output NHWC-p."""

    @staticmethod
    def forward(ctx, img, weight, bias, scale):
        require_cuda(img, 'from_rgb')
        b, cin, h, w = img.shape
        cout = weight.shape[0]
        pout = pitch_of(cout)
        dev = img.device
        w2 = weight.detach().reshape(cout, cin).contiguous()
        bias_c = bias.detach().contiguous() if bias is not None else None
        with torch.cuda.device(dev):
            y = _empty(b, h, w, pout, dev)
            sb, sc, sh, sw = img.stride()
            check(lib.cagc_from_rgb_fwd(stream_of(img), img.data_ptr(), sb, sc, sh, sw, w2.data_ptr(), ptr(bias_c),
                                        y.data_ptr(), b, h, w, cin, cout, pout, scale, 1, SQRT2), 'from_rgb_fwd')
        ctx.save_for_backward(y, w2)
        ctx.cfg = (b, cin, cout, h, w, pout, scale)
        return nhwc_view(y, cout)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        y, w2 = ctx.saved_tensors
        b, cin, cout, h, w, pout, scale = ctx.cfg
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        dev = g.device
        with torch.cuda.device(dev):
            gb = as_nhwc_buf(g)
            gimg = torch.empty((b, cin, h, w), device=dev, dtype=torch.float32)
            check(lib.cagc_from_rgb_bwd(stream_of(g), gb.data_ptr(), y.data_ptr(), w2.data_ptr(), gimg.data_ptr(), b, h, w,
                                        cin, cout, pout, scale, 1, SQRT2), 'from_rgb_bwd')
        return gimg, None, None, None


def from_rgb(img, conv, act_module):
    return _FromRGBFn.apply(img, conv.weight, act_module.bias, conv.scale)
