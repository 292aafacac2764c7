"""Autograd functions of the fused modulated-convolution path (host side of the C ABI).

Fused form of model.py:241-289 / :351-367 / :380-395 (math: SURVEY.md Appendix B):

    s_l = modulation_l(latent)   for ALL layers     <- ONE native kernel (style_affine), one more for its backward
    prep(W)  -> operand slabs (fwd + dgrad), Wsq, padded bias   <- ONE native kernel per layer (cached when frozen)
    d = rsqrt(s^2 @ Wsq^T + eps)                                <- ONE native kernel (demod)
    a = lrelu(d * conv(s*x, c*W) + nw*noise + bias) * sqrt2      <- modulate + ONE conv kernel (same-res)
    a = lrelu(d * blur(convT2(s*x, c*W)) + nw*noise + bias)*sqrt2 <- modulate + conv (all 4 parities) + FIR (up-conv)

The backward of a layer is act_bwd -> finalize (g_bias, gq = dL/d(demod radicand), g_noise_w) -> [FIR] -> dgrad
conv -> mod_bwd -> style-gradient finalize -> wgrad (split partials) -> wgrad finalize (parameter layout, demod
term): the demodulation terms of dW and ds are folded into the finalize kernels, nothing is left to ATen autograd.

Activations travel between layers as "NHWC-p" buffers: fp32 [B, H, W, P] with the channel pitch P
rounded up to a multiple of 8 and the padding channels kept at zero, exposed to Python as ordinary
[B, C, H, W] tensors (strided views), so the module surface stays that of the reference.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from typing import Optional

import torch
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ._lib import lib, check, stream_of, require_cuda, ptr, fir_nhwc, TensorCache, conv_workspace
from . import config


import os as _os

_FOLD_MOD = _os.environ.get('CAGC_FOLD_MOD', '1') != '0'      # development knob: modulation folded into per-sample weights


def pitch_of(c: int) -> int:
    return (c + 7) // 8 * 8


def nhwc_view(buf: torch.Tensor, c: int) -> torch.Tensor:
    """[B,H,W,P] buffer -> logical [B,C,H,W] view."""
    return buf[..., :c].permute(0, 3, 1, 2)


def as_nhwc_buf(x: torch.Tensor) -> torch.Tensor:
    """Logical [B,C,H,W] tensor -> NHWC-p buffer [B,H,W,P]; zero-copy when x already is one."""
    b, c, h, w = x.shape
    p = pitch_of(c)
    if (x.stride() == (h * w * p, 1, w * p, p) or (b == 1 and x.stride()[1:] == (1, w * p, p))) \
            and x.storage_offset() % 4 == 0 and x.data_ptr() % 16 == 0 \
            and x.untyped_storage().nbytes() >= (x.storage_offset() + b * h * w * p) * 4:
        return x.as_strided((b, h, w, p), (h * w * p, w * p, p, 1))
    buf = torch.empty((b, h, w, p), device=x.device, dtype=torch.float32)
    if x.numel():
        sb, sc, sh, sw = x.stride()
        check(lib.cagc_to_nhwc(stream_of(x), x.data_ptr(), sb, sc, sh, sw, buf.data_ptr(), b, c, h, w, p), 'to_nhwc')
    else:
        buf.zero_()
    return buf


_frozen = TensorCache()


def cached_frozen(param: torch.Tensor, tag, fn):
    """Weight-derived tensors (scaled weights, operand slabs, sum-of-squares for the demodulation) of
    a FROZEN parameter -- the teacher generator in the KD step -- are computed once and reused; a
    trainable parameter is always re-derived (its storage may be updated behind torch's version
    counter by the fused optimizer kernel)."""
    if param.requires_grad or torch.is_grad_enabled() and param.grad_fn is not None:
        return fn()

    def make():
        with torch.no_grad():
            return fn()
    return _frozen.get(param, tag, make)


def _timed(name, flops, nbytes, fn, shape=None):
    """Run a native launch, bracketing it with CUDA events when bench.py installed a profiler.  `shape`
    additionally files the launch under a per-shape key (the roofline of the dominant LAYER, next to the
    aggregate over all launches of the kernel)."""
    prof = config.profiler()
    if prof is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    prof.add(name, flops, nbytes, e0, e1)
    if shape is not None:
        prof.add(f'{name}@{shape}', flops, nbytes, e0, e1)
    return r


def _pad_last(t: torch.Tensor, p: int) -> torch.Tensor:
    t = t.contiguous()
    if t.shape[-1] == p:
        return t
    return F.pad(t, (0, p - t.shape[-1]))


def _use_tc(algo: int, k_pitch: int) -> bool:
    """The tcgen05 kernel consumes K in 128-byte (32-channel) TMA boxes; narrower layers (a few
    channels, latency-bound anyway) stay on the SIMT engine."""
    return algo == config.ALGO_TCGEN05_TF32 and k_pitch >= 32


def _r16(n: int) -> int:
    return (n + 15) // 16 * 16


def _is_frozen(t: Optional[torch.Tensor]) -> bool:
    return t is None or not (t.requires_grad or t.grad_fn is not None)


class _Prep:
    """Weight-derived operands of one modulated convolution (one cagc_weight_prep launch)."""
    __slots__ = ('w_fwd', 'w_dgrad', 'wsq_oi', 'wsq_io', 'bias_p')


def _weight_prep(weight, bias, wscale, upsample, tc_fwd, tc_dgrad, pin, pout, want_wsq):
    """weight [1,O,I,k,k] -> forward / data-gradient operand slabs in the layout of the engine that
    will consume them, Wsq in both orientations, zero-padded bias.  One launch when forward and
    data gradient run on the same engine (the usual case), two otherwise."""
    _, cout, cin, k, _ = weight.shape
    dev = weight.device
    kk = k * k
    r = _Prep()
    w = weight.detach()
    if not w.is_contiguous():
        w = w.contiguous()
    st = stream_of(w)
    # forward: tensor pipe wants K-major rows [t][o][i] (orientation A); SIMT wants [t][i][o] (orientation B)
    # dgrad  : tensor pipe [t'][i][o] (B), SIMT [t'][o][i] (A); taps flipped for the same-resolution conv
    flip_d = 0 if upsample else 1
    f_spec = ('A', _r16(pout), pin, 0, 1) if tc_fwd else ('B', pin, pout, 0, 0)
    d_spec = ('B', _r16(pin), pout, flip_d, 1) if tc_dgrad else ('A', pout, pin, flip_d, 0)
    r.w_fwd = torch.empty((kk, f_spec[1], f_spec[2]), device=dev, dtype=torch.float32)
    r.w_dgrad = torch.empty((kk, d_spec[1], d_spec[2]), device=dev, dtype=torch.float32)
    if want_wsq:
        r.wsq_oi = torch.empty((pout, pin), device=dev, dtype=torch.float32)
        r.wsq_io = torch.empty((pin, pout), device=dev, dtype=torch.float32)
    else:
        r.wsq_oi = r.wsq_io = None
    if bias is not None:
        b_in = bias.detach().contiguous()
        r.bias_p = torch.empty((pout,), device=dev, dtype=torch.float32)
    else:
        b_in, r.bias_p = None, None

    def call(spec_a, out_a, spec_b, out_b, rnd, with_rest):
        ra, ca, fa = (spec_a[1], spec_a[2], spec_a[3]) if spec_a else (0, 0, 0)
        rb, cb, fb = (spec_b[1], spec_b[2], spec_b[3]) if spec_b else (0, 0, 0)
        check(lib.cagc_weight_prep(st, w.data_ptr(), wscale, cout, cin, k,
                                   ptr(out_a), ra, ca, fa, ptr(out_b), rb, cb, fb, rnd,
                                   ptr(r.wsq_oi) if with_rest else None, ptr(r.wsq_io) if with_rest else None,
                                   pout, pin, ptr(b_in) if with_rest else None,
                                   ptr(r.bias_p) if with_rest else None, cout, pout), 'weight_prep')

    if f_spec[0] != d_spec[0] and f_spec[4] == d_spec[4]:
        a, b = (f_spec, d_spec) if f_spec[0] == 'A' else (d_spec, f_spec)
        oa, ob = (r.w_fwd, r.w_dgrad) if f_spec[0] == 'A' else (r.w_dgrad, r.w_fwd)
        call(a, oa, b, ob, f_spec[4], True)
    else:
        for spec, out, rest in ((f_spec, r.w_fwd, True), (d_spec, r.w_dgrad, False)):
            if spec[0] == 'A':
                call(spec, out, None, None, spec[4], rest)
            else:
                call(None, None, spec, out, spec[4], rest)
    return r


def _prepared(weight, bias, wscale, upsample, tc_fwd, tc_dgrad, pin, pout, want_wsq):
    """Frozen parameters (the teacher generator in the KD step): prepared once and reused.  A trainable
    parameter is re-derived every call (its storage may be updated behind torch's version counter by
    the fused optimizer kernel)."""
    if not (_is_frozen(weight) and _is_frozen(bias)):
        return _weight_prep(weight, bias, wscale, upsample, tc_fwd, tc_dgrad, pin, pout, want_wsq)
    tag = ('prep', wscale, upsample, tc_fwd, tc_dgrad, want_wsq, None if bias is None else (id(bias), bias._version))
    hit = _frozen.get(weight, tag, lambda: (_weight_prep(weight, bias, wscale, upsample, tc_fwd, tc_dgrad, pin, pout,
                                                        want_wsq), None if bias is None else weakref.ref(bias)))
    if hit[1] is not None and hit[1]() is not bias:      # the bias object that shared this id is gone: rebuild
        return _weight_prep(weight, bias, wscale, upsample, tc_fwd, tc_dgrad, pin, pout, want_wsq)
    return hit[0]


class _StyledConvFn(Function):
    """a = [lrelu]( d * conv(s*x, c*W) [+ nw*noise] [+ bias] ) on NHWC-p buffers, d = rsqrt(s^2 Wsq + eps).

    Inputs: x [B,I,H,W]; s_p [B,pitch(I)] (zero padded); weight [1,O,I,k,k]; noise [B|1,1,Ho,Wo] or None;
    noise_w [1] or None; bias [O] or None.  Non-tensor: wscale (c), demodulate, eps, upsample, fir [4,4], pad, act.
    """

    @staticmethod
    def forward(ctx, x, s_p, weight, noise, noise_w, bias, wscale, demodulate, eps, upsample, fir, pad, act, algo):
        require_cuda(x, 'ModulatedConv2d')
        b, cin, h, w = x.shape
        _, cout, cin_w, k, _ = weight.shape
        if cin_w != cin:
            raise RuntimeError(f'ModulatedConv2d: input has {cin} channels, weight expects {cin_w}')
        pin, pout = pitch_of(cin), pitch_of(cout)
        if s_p.shape != (b, pin):
            raise RuntimeError(f'ModulatedConv2d: style scalars have shape {tuple(s_p.shape)}, expected {(b, pin)}')
        dev = x.device
        with torch.cuda.device(dev):
            st = stream_of(x)
            xb = as_nhwc_buf(x.detach())
            s_p = s_p.detach().contiguous()
            tc = _use_tc(algo, pin)
            tc_d = _use_tc(algo, pout)
            prep = _prepared(weight, bias, wscale, upsample, tc, tc_d, pin, pout, demodulate)
            bias_p = prep.bias_p
            if demodulate:
                d_p = torch.empty((b, pout), device=dev, dtype=torch.float32)
                check(lib.cagc_demod(st, s_p.data_ptr(), prep.wsq_io.data_ptr(), d_p.data_ptr(), b, pin, cout, pout,
                                     eps), 'demod')
            else:
                d_p = None
            # A layer nobody differentiates (the frozen teacher under no_grad; sampling) does not need the modulated
            # activation as a weight-gradient operand: where the halo-tile kernel takes the shape, the modulation is
            # folded into per-sample weight slabs (B x k^2 small slabs) and the activation is read as it is.
            psw_bytes = 0
            if tc and not upsample and _FOLD_MOD and xb.numel() and not any(ctx.needs_input_grad):
                psw_bytes = int(lib.cagc_conv_same_psw_bytes(b, h, w, pin, pout, k))
            if tc and psw_bytes:
                x_in, s_arg, falgo = None, None, config.ALGO_TCGEN05_TF32
            elif tc:
                # tensor-pipe path: operands come straight from TMA, so modulation is a tensor pass
                x_in = torch.empty_like(xb)
                if xb.numel():
                    check(lib.cagc_modulate(st, xb.data_ptr(), s_p.data_ptr(), x_in.data_ptr(), b, h, w, pin), 'modulate')
                s_arg, falgo = None, config.ALGO_TCGEN05_TF32
            else:
                x_in, s_arg, falgo = xb, s_p.data_ptr(), config.ALGO_SIMT_FP32
            w_fwd = prep.w_fwd
            if noise is not None:
                require_cuda(noise, 'ModulatedConv2d(noise)')
                if noise.device != dev:
                    raise RuntimeError(f'noise lives on {noise.device}, the activation on {dev}')
                noise = noise.detach().contiguous()
                nb = noise.shape[0]
                ho, wo = (2 * h, 2 * w) if upsample else (h, w)
                if noise.shape[1:] != (1, ho, wo) or nb not in (1, b):
                    raise RuntimeError(f'noise shape {tuple(noise.shape)} does not match output {(b, 1, ho, wo)}')
                nstride = 0 if (nb == 1 and b != 1) else ho * wo
                nw = noise_w.detach().reshape(1).contiguous()
            else:
                nstride, nw = 0, None
            if not upsample:
                out = torch.empty((b, h, w, pout), device=dev, dtype=torch.float32)
                if out.numel() and psw_bytes:
                    w_ps = torch.empty(psw_bytes // 4, device=dev, dtype=torch.float32)
                    _timed(f'conv_same[algo{falgo}]', 2.0 * b * h * w * cin * cout * k * k, 4.0 * b * h * w * (cin + cout),
                           lambda: check(lib.cagc_conv_same_psw(st, xb.data_ptr(), w_fwd.data_ptr(), s_p.data_ptr(),
                                                                ptr(d_p), ptr(noise), ptr(nw), ptr(bias_p), out.data_ptr(),
                                                                b, h, w, pin, pout, cout, k, nstride, int(act),
                                                                w_ps.data_ptr(), psw_bytes),
                                         'conv_same_psw'), shape=f'{cin}->{cout}x{h}x{w}')
                    del w_ps
                elif out.numel():
                    flops = 2.0 * b * h * w * cin * cout * k * k          # Util/Calculators.py convention x2
                    ws, ws_bytes = conv_workspace(b, h, w, pout, dev) if tc else (None, 0)
                    _timed(f'conv_same[algo{falgo}]', flops, 4.0 * b * h * w * (cin + cout),
                           lambda: check(lib.cagc_conv_same_ws(st, x_in.data_ptr(), w_fwd.data_ptr(), s_arg,
                                                               ptr(d_p), ptr(noise), ptr(nw), ptr(bias_p), out.data_ptr(),
                                                               b, h, w, pin, pout, cout, k, nstride, int(act), falgo,
                                                               ptr(ws), ws_bytes),
                                         'conv_same'), shape=f'{cin}->{cout}x{h}x{w}')
                    del ws
            else:
                hu, wu = 2 * h + k - 2, 2 * w + k - 2
                ut = torch.empty((b, hu, wu, pout), device=dev, dtype=torch.float32)
                fir = fir.contiguous()
                kh, kw = fir.shape
                ho, wo = hu + pad[0] + pad[1] - kh + 1, wu + pad[0] + pad[1] - kw + 1
                out = torch.empty((b, ho, wo, pout), device=dev, dtype=torch.float32)
                if out.numel():
                    flops = 2.0 * b * h * w * cin * cout * k * k          # transposed conv counted at input res
                    ws, ws_bytes = conv_workspace(b, hu, wu, pout, dev) if tc else (None, 0)
                    _timed(f'conv_up[algo{falgo}]', flops, 4.0 * b * (h * w * cin + hu * wu * cout),
                           lambda: check(lib.cagc_conv_up_ws(st, x_in.data_ptr(), w_fwd.data_ptr(), s_arg,
                                                             ut.data_ptr(), b, h, w, pin, pout, k, falgo, ptr(ws), ws_bytes),
                                         'conv_up'),
                           shape=f'{cin}->{cout}x{h}x{w}')
                    del ws
                    _timed('fir_nhwc', 0.0, 4.0 * b * cout * (hu * wu + ho * wo),
                           lambda: fir_nhwc(st, ut.data_ptr(), fir, d_p, noise, nw, bias_p, out.data_ptr(), b, hu, wu,
                                            pout, cout, (pad[0], pad[1], pad[0], pad[1]), nstride, int(act), 'fir_nhwc'),
                           shape=f'{cout}x{hu}x{wu}')
                del ut
        xm = x_in if tc else None      # modulated, TF32-rounded input: A operand of the tensor-pipe wgrad
        ctx.save_for_backward(xb, s_p, d_p, weight, noise, nw, bias_p, out, fir if upsample else None, xm,
                              prep.w_dgrad, prep.wsq_oi)
        ctx.cfg = (b, cin, cout, h, w, k, pin, pout, upsample, pad, bool(act), nstride, algo, wscale,
                   demodulate, bias is not None)
        return nhwc_view(out, cout)

    @staticmethod
    @once_differentiable
    def backward(ctx, ga):
        xb, s_p, d_p, weight, noise, nw, bias_p, out, fir, xm, w_d, wsq_oi = ctx.saved_tensors
        (b, cin, cout, h, w, k, pin, pout, upsample, pad, act, nstride, algo, wscale, has_d, has_bias) = ctx.cfg
        dev = xb.device
        ho, wo = out.shape[1], out.shape[2]
        need_x, need_s, need_w = ctx.needs_input_grad[0:3]
        need_nw, need_bias = ctx.needs_input_grad[4], ctx.needs_input_grad[5]
        with torch.cuda.device(dev):
            st = stream_of(xb)
            gu = torch.empty_like(out)
            chunks = lib.cagc_act_bwd_chunks(ho, wo)
            partial = torch.empty((b, chunks, 3, pout), device=dev, dtype=torch.float32)
            sb, sc, sh, sw = ga.stride()
            check(lib.cagc_act_bwd(st, ga.data_ptr(), sb, sc, sh, sw, out.data_ptr(), ptr(d_p), ptr(noise), ptr(nw),
                                   ptr(bias_p), gu.data_ptr(), partial.data_ptr(), b, ho, wo, pout, cout,
                                   nstride, int(act)), 'act_bwd')
            # chunk partials -> g_bias, gq = dL/d(demodulation radicand) = -1/2 d^3 gd, noise-weight partials
            want_bias = has_bias and need_bias
            want_q = has_d and (need_s or need_w)
            want_nw = noise is not None and need_nw
            g_bias_p = torch.empty((pout,), device=dev, dtype=torch.float32) if want_bias else None
            gq = torch.empty((b, pout), device=dev, dtype=torch.float32) if want_q else None
            nblk = lib.cagc_act_bwd_finalize_blocks(pout)
            nw_part = torch.empty((nblk,), device=dev, dtype=torch.float32) if want_nw else None
            if want_bias or want_q or want_nw:
                check(lib.cagc_act_bwd_finalize(st, partial.data_ptr(), ptr(d_p), ptr(g_bias_p), ptr(gq), ptr(nw_part),
                                                b, chunks, pout), 'act_bwd_finalize')
            g_bias = g_bias_p[:cout] if want_bias else None
            g_nw = nw_part.sum().reshape(1) if want_nw else None
            del partial

            if upsample:
                # blur backward: correlation with the FIR kernel, pads of op/upfirdn2d.py:111-116
                kh, kw = fir.shape
                hu, wu = 2 * h + k - 2, 2 * w + k - 2
                gp0, gp1 = kh - pad[0] - 1, hu - ho + pad[0]
                g_t = torch.empty((b, hu, wu, pout), device=dev, dtype=torch.float32)
                firf = _flipped(fir)
                _timed('fir_nhwc', 0.0, 4.0 * b * cout * (hu * wu + ho * wo),
                       lambda: fir_nhwc(st, gu.data_ptr(), firf, None, None, None, None, g_t.data_ptr(), b, ho, wo, pout,
                                        pout, (gp0, gp1, gp0, gp1), 0, 0, 'fir_nhwc(bwd)'))
                g_conv = g_t
            else:
                g_conv = gu

            g_x = g_s = g_w = None
            if need_x or need_s:
                gxt = torch.empty((b, h, w, pin), device=dev, dtype=torch.float32)
                dalgo = config.ALGO_TCGEN05_TF32 if _use_tc(algo, pout) else config.ALGO_SIMT_FP32
                flops, nbytes = 2.0 * b * h * w * cin * cout * k * k, 4.0 * b * h * w * (cin + cout)
                ws, ws_bytes = conv_workspace(b, h, w, pin, dev) if dalgo == config.ALGO_TCGEN05_TF32 else (None, 0)
                if upsample:
                    _timed(f'conv_up_dgrad[algo{dalgo}]', flops, 4.0 * b * (h * w * cin + hu * wu * cout),
                           lambda: check(lib.cagc_conv_up_dgrad_ws(st, g_conv.data_ptr(), w_d.data_ptr(), gxt.data_ptr(),
                                                                   b, h, w, pout, pin, k, dalgo, ptr(ws), ws_bytes),
                                         'conv_up_dgrad'))
                else:
                    _timed(f'conv_same[algo{dalgo}]', flops, nbytes,
                           lambda: check(lib.cagc_conv_same_ws(st, g_conv.data_ptr(), w_d.data_ptr(), None, None, None,
                                                               None, None, gxt.data_ptr(), b, h, w, pout, pin, pin, k, 0,
                                                               0, dalgo, ptr(ws), ws_bytes), 'conv_same(dgrad)'))
                del ws
                mchunks = lib.cagc_act_bwd_chunks(h, w)
                mpartial = torch.empty((b, mchunks, pin), device=dev, dtype=torch.float32)
                check(lib.cagc_mod_bwd(st, gxt.data_ptr(), xb.data_ptr(), s_p.data_ptr(), mpartial.data_ptr(),
                                       b, h, w, pin), 'mod_bwd')
                if need_s:
                    g_s = torch.empty((b, pin), device=dev, dtype=torch.float32)
                    check(lib.cagc_style_grad_finalize(st, mpartial.data_ptr(), ptr(gq), s_p.data_ptr(), ptr(wsq_oi),
                                                       g_s.data_ptr(), b, mchunks, pin, cout, pout),
                          'style_grad_finalize')
                if need_x:
                    g_x = nhwc_view(gxt, cin)
            if need_w:
                # exact-fp32 mode (saliency): SIMT engine; tensor-pipe mode: tcgen05 TF32.  Both leave split-K
                # partials that cagc_wgrad_finalize reduces in a fixed order (deterministic), folding in the
                # weight scale, the demodulation term and the [O,I,k,k] parameter layout
                walgo = config.ALGO_TCGEN05_TF32 if (xm is not None and _use_tc(algo, min(pin, pout))
                                                     and pout <= 256) else config.ALGO_SIMT_FP32
                nsp = lib.cagc_conv_wgrad_splits(b, h, w, pin, pout, k, walgo)
                wpart = torch.empty((nsp, k * k, pin, pout), device=dev, dtype=torch.float32)
                a_in, a_sc = (xm, None) if walgo == config.ALGO_TCGEN05_TF32 else (xb, s_p.data_ptr())
                used = C.c_int(0)
                _timed(f'conv_wgrad[algo{walgo}]', 2.0 * b * h * w * cin * cout * k * k,
                       4.0 * b * h * w * (cin + cout),
                       lambda: check(lib.cagc_conv_wgrad_partial(st, a_in.data_ptr(), a_sc, g_conv.data_ptr(),
                                                                 wpart.data_ptr(), nsp, b, h, w, pin, pout, k,
                                                                 1 if upsample else 0, walgo, C.byref(used)),
                                     'conv_wgrad'))
                g_w = torch.empty((1, cout, cin, k, k), device=dev, dtype=torch.float32)
                wc = weight.detach()
                wc = wc if wc.is_contiguous() else wc.contiguous()
                check(lib.cagc_wgrad_finalize(st, wpart.data_ptr(), used.value, wc.data_ptr(), wscale, ptr(gq),
                                              s_p.data_ptr(), b, cout, cin, k, pin, pout, g_w.data_ptr()),
                      'wgrad_finalize')
        return (g_x, g_s, g_w, None, g_nw, g_bias, None, None, None, None, None, None, None, None)


def _flipped(fir: torch.Tensor) -> torch.Tensor:
    """flip(fir) of a (constant) FIR buffer, computed once per buffer version."""
    return _frozen.get(fir, 'flip', lambda: torch.flip(fir.detach(), [0, 1]).contiguous())


def styled_conv(x, s_p, weight, noise, noise_w, bias, wscale, demodulate=True, eps=1e-8, upsample=False, fir=None,
                pad=(0, 0), act=True):
    return _StyledConvFn.apply(x, s_p, weight, noise, noise_w, bias, wscale, demodulate, eps, upsample, fir, pad, act,
                               config.conv_algo())


class _ToRGBFn(Function):
    """rgb = conv1x1(s*x, c*W) + bias + Upsample(skip), output NCHW [B,3,H,W] (model.py:380-395)."""

    @staticmethod
    def forward(ctx, x, s_p, weight, bias, skip, wscale, fir, pad, passthrough=False):
        require_cuda(x, 'ToRGB')
        b, cin, h, w = x.shape
        nout = weight.shape[1]
        pin = pitch_of(cin)
        if s_p.shape != (b, pin):
            raise RuntimeError(f'ToRGB: style scalars have shape {tuple(s_p.shape)}, expected {(b, pin)}')
        dev = x.device
        with torch.cuda.device(dev):
            st = stream_of(x)
            xb = as_nhwc_buf(x.detach())
            s_p = s_p.detach().contiguous()
            w2 = weight.detach().reshape(nout, cin).contiguous()
            bias_c = bias.detach().reshape(nout).contiguous() if bias is not None else None
            out = torch.empty((b, nout, h, w), device=dev, dtype=torch.float32)
            if skip is not None:
                skip_c = skip.detach().contiguous()
                if skip_c.shape != (b, nout, h // 2, w // 2):
                    raise RuntimeError(f'ToRGB: skip shape {tuple(skip_c.shape)} does not match {(b, nout, h // 2, w // 2)}')
                fir = fir.contiguous()
                fh, fw = fir.shape
            else:
                skip_c, fh, fw = None, 0, 0
            if out.numel():
                check(lib.cagc_torgb_fwd(st, xb.data_ptr(), w2.data_ptr(), s_p.data_ptr(), ptr(bias_c), ptr(skip_c),
                                         ptr(fir) if skip is not None else None, out.data_ptr(), b, h, w, pin, cin, nout,
                                         wscale, fh, fw, pad[0], pad[1]), 'torgb_fwd')
        ctx.save_for_backward(xb, s_p, w2, fir if skip is not None else None)
        ctx.cfg = (b, cin, nout, h, w, pin, wscale, pad, bias is not None, skip is not None)
        if passthrough:
            # the activation is handed on to its other consumer THROUGH this node: both of its gradients then
            # arrive here and are summed inside torgb_bwd instead of by a separate accumulation kernel
            return out, x
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g, g_through=None):
        xb, s_p, w2, fir = ctx.saved_tensors
        b, cin, nout, h, w, pin, wscale, pad, has_bias, has_skip = ctx.cfg
        dev = xb.device
        if g is None:       # only the pass-through branch reached the loss
            return g_through, None, None, None, None, None, None, None, None
        g = g.contiguous()
        with torch.cuda.device(dev):
            st = stream_of(xb)
            gx = torch.empty((b, h, w, pin), device=dev, dtype=torch.float32)
            chunks = lib.cagc_act_bwd_chunks(h, w)
            partial = torch.empty((b, chunks, nout, pin), device=dev, dtype=torch.float32)
            g_add = as_nhwc_buf(g_through) if g_through is not None else None
            check(lib.cagc_torgb_bwd(st, g.data_ptr(), xb.data_ptr(), w2.data_ptr(), s_p.data_ptr(), ptr(g_add),
                                     gx.data_ptr(), partial.data_ptr(), b, h, w, pin, cin, nout, wscale), 'torgb_bwd')
            g_w = g_s = g_bias = g_skip = None
            need_s, need_w = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
            if need_s or need_w:
                gw_part = torch.empty((b, nout, cin), device=dev, dtype=torch.float32) if need_w else None
                g_s = torch.empty((b, pin), device=dev, dtype=torch.float32) if need_s else None
                check(lib.cagc_torgb_bwd_finalize(st, partial.data_ptr(), s_p.data_ptr(), w2.data_ptr(), wscale,
                                                  ptr(gw_part), ptr(g_s), b, chunks, cin, pin, nout), 'torgb_bwd_finalize')
                g_w = gw_part.sum(0).reshape(1, nout, cin, 1, 1) if need_w else None   # fixed-order sum over the batch
            if has_bias and ctx.needs_input_grad[3]:
                g_bias = g.sum(dim=(0, 2, 3)).reshape(1, nout, 1, 1)
            if has_skip and ctx.needs_input_grad[4]:
                from op.upfirdn2d import _launch
                fh, fw = fir.shape
                # backward of Upsample(up=2, pad): down=2 with the flipped kernel (op/upfirdn2d.py:111-116)
                gp0 = fh - pad[0] - 1
                gp1 = (h // 2) * 2 - h + pad[0] - 2 + 1
                g_skip = _launch(g, _flipped(fir), (1, 1), (2, 2), (gp0, gp1, gp0, gp1))
            g_x = nhwc_view(gx, cin) if ctx.needs_input_grad[0] else None
        return g_x, g_s, g_w, g_bias, g_skip, None, None, None, None


def to_rgb(x, s_p, weight, bias, skip, wscale, fir=None, pad=(0, 0), passthrough=False):
    """passthrough=True returns (rgb, x'): x' aliases x and carries its gradient back through this node."""
    return _ToRGBFn.apply(x, s_p, weight, bias, skip, wscale, fir, pad, passthrough)


# ------------------------------------------------------------------------------------------------
# Style modulation of many layers in one launch (EqualLinear, model.py:156-166, as used at :248)
# ------------------------------------------------------------------------------------------------
def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


class _StyleAffineFn(Function):
    """s_l = scale_l * latent[:, idx_l] @ A_l^T + lr_mul_l * bias_l for every layer l of `meta`, zero padded
    to the channel pitch.  args: latent [B, n_latent, D], meta = [(I, pitch, idx, scale, lr_mul)], then
    A_0, bias_0, A_1, bias_1, ...  Returns one [B, pitch_l] tensor per layer."""

    @staticmethod
    def forward(ctx, latent, meta, *params):
        require_cuda(latent, 'style modulation')
        n = len(meta)
        b, n_latent, dim = latent.shape
        lat = latent.detach()
        if lat.stride(2) != 1:
            lat = lat.contiguous()
        As = [params[2 * i].detach().contiguous() for i in range(n)]
        bs = [None if params[2 * i + 1] is None else params[2 * i + 1].detach().contiguous() for i in range(n)]
        dev = latent.device
        outs = [torch.empty((b, m[1]), device=dev, dtype=torch.float32) for m in meta]
        arr_i = (C.c_int * n)(*[m[0] for m in meta])
        arr_p = (C.c_int * n)(*[m[1] for m in meta])
        arr_l = (C.c_int * n)(*[m[2] for m in meta])
        arr_s = (C.c_float * n)(*[m[3] for m in meta])
        arr_m = (C.c_float * n)(*[m[4] for m in meta])
        with torch.cuda.device(dev):
            check(lib.cagc_style_affine(stream_of(lat), n, _ptr_array(As), _ptr_array(bs), _ptr_array(outs), arr_i,
                                        arr_p, arr_l, arr_s, arr_m, lat.data_ptr(), lat.stride(0), lat.stride(1),
                                        b, dim, n_latent), 'style_affine')
        ctx.save_for_backward(lat, *As)
        ctx.meta = meta
        ctx.has_bias = [x is not None for x in bs]
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *gs):
        lat, *As = ctx.saved_tensors
        meta = ctx.meta
        n = len(meta)
        b, n_latent, dim = lat.shape
        dev = lat.device
        gs = [None if g is None else g.contiguous() for g in gs]
        need_lat = ctx.needs_input_grad[0]
        need_par = any(ctx.needs_input_grad[2:])
        gA = [torch.empty_like(a) for a in As] if need_par else None
        gb = [torch.empty((m[0],), device=dev, dtype=torch.float32) if hb else None
              for m, hb in zip(meta, ctx.has_bias)] if need_par else None
        g_lat = torch.empty((b, n_latent, dim), device=dev, dtype=torch.float32) if need_lat else None
        arr_i = (C.c_int * n)(*[m[0] for m in meta])
        arr_p = (C.c_int * n)(*[m[1] for m in meta])
        arr_l = (C.c_int * n)(*[m[2] for m in meta])
        arr_s = (C.c_float * n)(*[m[3] for m in meta])
        arr_m = (C.c_float * n)(*[m[4] for m in meta])
        with torch.cuda.device(dev):
            check(lib.cagc_style_affine_bwd(stream_of(lat), n, _ptr_array(As), _ptr_array(gs),
                                            _ptr_array(gA) if need_par else None,
                                            _ptr_array(gb) if need_par else None, arr_i, arr_p, arr_l, arr_s, arr_m,
                                            lat.data_ptr(), lat.stride(0), lat.stride(1), ptr(g_lat), b, dim, n_latent),
                  'style_affine_bwd')
        grads = [g_lat, None]
        for i in range(n):
            grads.append(gA[i] if need_par and ctx.needs_input_grad[2 + 2 * i] else None)
            grads.append(gb[i] if need_par and ctx.has_bias[i] and ctx.needs_input_grad[3 + 2 * i] else None)
        return tuple(grads)


_MAX_STYLE_LAYERS = 40


def style_affine(latent: torch.Tensor, mods, indices):
    """latent [B, n_latent, D]; mods: EqualLinear-like objects (weight [I,D], bias [I] or None, scale,
    lr_mul); indices: which latent row each layer reads.  Returns the list of zero-padded style
    scalars s_l [B, pitch(I_l)] -- one launch for the whole generator."""
    outs = []
    for lo in range(0, len(mods), _MAX_STYLE_LAYERS):
        chunk = mods[lo:lo + _MAX_STYLE_LAYERS]
        meta = [(m.weight.shape[0], pitch_of(m.weight.shape[0]), int(i), float(m.scale), float(m.lr_mul))
                for m, i in zip(chunk, indices[lo:lo + _MAX_STYLE_LAYERS])]
        params = []
        for m in chunk:
            params += [m.weight, m.bias]
        outs += list(_StyleAffineFn.apply(latent, meta, *params))
    return outs


# ------------------------------------------------------------------------------------------------
# EqualLinear (model.py:137-171): own skinny-GEMM kernels with the epilogue fused (csrc/linear.cu)
# ------------------------------------------------------------------------------------------------
class _EqualLinearFn(Function):
    """y = [lrelu * sqrt2](scale * x W^T + bias * lr_mul), forward = ONE native launch (cagc_linear_fwd), backward =
    activation/bias kernel + the two small products (cagc_linear_bwd); no library GEMM."""

    @staticmethod
    def forward(ctx, x, weight, bias, scale, lr_mul, act):
        require_cuda(x, 'EqualLinear')
        n_out, n_in = weight.shape
        x2 = x.detach().reshape(-1, n_in)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        w = weight.detach()
        if not w.is_contiguous():
            w = w.contiguous()
        m = x2.shape[0]
        out = torch.empty((m, n_out), device=x.device, dtype=torch.float32)
        bias_c = bias.detach().contiguous() if bias is not None else None
        with torch.cuda.device(x.device):
            check(lib.cagc_linear_fwd(stream_of(x2), x2.data_ptr(), w.data_ptr(), ptr(bias_c), out.data_ptr(), m, n_out,
                                      n_in, scale, lr_mul, int(act), 0.2, math.sqrt(2)), 'linear_fwd')
        ctx.save_for_backward(x2, w, out if act else None)
        ctx.cfg = (scale, lr_mul, act, bias is not None, x.shape)
        ctx.weight_param = weight        # identity key of the transposed copy the backward caches for frozen weights
        return out.reshape(*x.shape[:-1], n_out)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x2, w, out = ctx.saved_tensors
        scale, lr_mul, act, has_bias, xshape = ctx.cfg
        n_out, n_in = w.shape
        m = x2.shape[0]
        g2 = g.reshape(m, n_out).contiguous()
        g_acc = torch.empty_like(g2)
        g_bias = torch.empty((n_out,), device=g.device, dtype=torch.float32) \
            if (has_bias and ctx.needs_input_grad[2]) else None
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_x = torch.empty((m, n_in), device=g.device, dtype=torch.float32) if need_x else None
        g_w = torch.empty((n_out, n_in), device=g.device, dtype=torch.float32) if need_w else None
        with torch.cuda.device(g.device):
            st = stream_of(g2)
            check(lib.cagc_linear_bias_act_bwd(st, g2.data_ptr(), ptr(out), g_acc.data_ptr(), ptr(g_bias),
                                               m, n_out, scale, lr_mul, int(act), 0.2, math.sqrt(2)),
                  'linear_bias_act_bwd')
            if need_x:
                # g_x = g_acc @ W is the forward kernel on W^T (one warp per output feature streaming a contiguous row);
                # the column-strided form (cagc_linear_bwd's g_x) left 16 blocks walking 512 dependent steps: 86 us
                wt = cached_frozen(ctx.weight_param, ('eqlin_t',), lambda: w.t().contiguous())
                check(lib.cagc_linear_fwd(st, g_acc.data_ptr(), wt.data_ptr(), None, g_x.data_ptr(), m, n_in, n_out,
                                          1.0, 0.0, 0, 0.2, 1.0), 'linear_bwd_x')
            if need_w:
                check(lib.cagc_linear_bwd(st, g_acc.data_ptr(), x2.data_ptr(), w.data_ptr(), None, ptr(g_w), m, n_out,
                                          n_in), 'linear_bwd_w')
        return (g_x.reshape(xshape) if need_x else None), g_w, g_bias, None, None, None


def equal_linear(x, weight, bias, scale, lr_mul, act):
    return _EqualLinearFn.apply(x, weight, bias, scale, lr_mul, bool(act))


def demod_coefficients(s: torch.Tensor, weight: torch.Tensor, wscale: float, eps: float = 1e-8) -> torch.Tensor:
    """d[b,o] = rsqrt(sum_i s[b,i]^2 * Wsq[o,i] + eps), Wsq = sum_taps (c*W)^2   (model.py:251-253);
    differentiable torch form, used by the second-order composite only."""
    wsq = (weight[0] * wscale).square().sum(dim=(2, 3))   # [O,I]
    return torch.rsqrt(s.square() @ wsq.t() + eps)


# ------------------------------------------------------------------------------------------------
# Differentiable composite (second-order path only: create_graph=True through the generator)
# ------------------------------------------------------------------------------------------------
def styled_conv_composite(x, s, d, weight, noise, noise_w, bias, wscale, upsample=False, fir=None, pad=(0, 0),
                          act=True):
    from op import upfirdn2d, fused_leaky_relu
    k = weight.shape[-1]
    xm = x * s[:, :, None, None]
    w = weight[0] * wscale
    if upsample:
        u = F.conv_transpose2d(xm, w.transpose(0, 1), stride=2)
        u = upfirdn2d(u.contiguous(), fir, pad=pad)
    else:
        u = F.conv2d(xm, w, padding=k // 2)
    if d is not None:
        u = u * d[:, :, None, None]
    if noise is not None:
        u = u + noise_w * noise
    if act:
        return fused_leaky_relu(u.contiguous(), bias)
    if bias is not None:
        u = u + bias.view(1, -1, 1, 1)
    return u


def to_rgb_composite(x, s, weight, bias, skip, wscale, fir=None, pad=(0, 0)):
    from op import upfirdn2d
    out = F.conv2d(x * s[:, :, None, None], weight[0] * wscale)
    if bias is not None:
        out = out + bias
    if skip is not None:
        out = out + upfirdn2d(skip, fir, up=2, down=1, pad=pad)
    return out
