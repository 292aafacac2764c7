"""Autograd functions of the fused modulated-convolution path (host side of the C ABI).

Fused form of model.py:241-289 / :351-367 / :380-395 (math: SURVEY.md Appendix B):

    s = modulation(style)                      -- EqualLinear, plain library GEMM, autograd by torch
    d = rsqrt(s^2 @ Wsq^T + eps)               -- tiny, composed from torch ops so autograd supplies
                                                  the demodulation terms of dW and ds
    a = lrelu(d * conv(s*x, c*W) + nw*noise + bias) * sqrt2      <- ONE native kernel (same-res)
    a = lrelu(d * blur(convT2(s*x, c*W)) + nw*noise + bias)*sqrt2 <- two native kernels (up-conv)

Activations travel between layers as "NHWC-p" buffers: fp32 [B, H, W, P] with the channel pitch P
rounded up to a multiple of 8 and the padding channels kept at zero, exposed to Python as ordinary
[B, C, H, W] tensors (strided views), so the module surface stays that of the reference.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch.autograd import Function

from ._lib import lib, check, stream_of, require_cuda, ptr
from . import config


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


_frozen = {}


def cached_frozen(param: torch.Tensor, tag, fn):
    """Weight-derived tensors (scaled weights, operand slabs, sum-of-squares for the demodulation) of
    a FROZEN parameter -- the teacher generator in the KD step -- are computed once and reused; a
    trainable parameter is always re-derived (its storage may be updated behind torch's version
    counter by the fused optimizer kernel)."""
    if param.requires_grad or torch.is_grad_enabled() and param.grad_fn is not None:
        return fn()
    key = (param.data_ptr(), param._version, tuple(param.shape), tag)
    hit = _frozen.get(key)
    if hit is None:
        if len(_frozen) > 4096:
            _frozen.clear()
        with torch.no_grad():
            hit = fn()
        _frozen[key] = hit
    return hit


def _timed(name, flops, nbytes, fn):
    """Run a native launch, bracketing it with CUDA events when bench.py installed a profiler."""
    prof = config.profiler()
    if prof is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    prof.add(name, flops, nbytes, e0, e1)
    return r


def _pad_last(t: torch.Tensor, p: int) -> torch.Tensor:
    t = t.contiguous()
    if t.shape[-1] == p:
        return t
    return F.pad(t, (0, p - t.shape[-1]))


def _slabs(w4: torch.Tensor, rows_p: int, cols_p: int) -> torch.Tensor:
    """[k,k,R,C] -> zero-padded contiguous [k*k, rows_p, cols_p]."""
    k = w4.shape[0]
    r, c = w4.shape[2], w4.shape[3]
    w = w4.reshape(k * k, r, c)
    if r == rows_p and c == cols_p:
        return w.contiguous()
    return F.pad(w, (0, cols_p - c, 0, rows_p - r)).contiguous()


def _slabs_tc(w4: torch.Tensor, k_p: int) -> torch.Tensor:
    """[k,k,N,K] -> K-major slabs [k*k, roundup16(pitch(N)), k_p], rounded to TF32 (tcgen05 B operand)."""
    k = w4.shape[0]
    n, kk = w4.shape[2], w4.shape[3]
    n_rows = (pitch_of(n) + 15) // 16 * 16
    w = F.pad(w4.reshape(k * k, n, kk), (0, k_p - kk, 0, n_rows - n)).contiguous()
    check(lib.cagc_modulate(stream_of(w), w.data_ptr(), None, w.data_ptr(), 1, 1, w.numel() // 4, 4), 'round_tf32')
    return w


def _use_tc(algo: int, k_pitch: int) -> bool:
    """The tcgen05 kernel consumes K in 128-byte (32-channel) TMA boxes; narrower layers (a few
    channels, latency-bound anyway) stay on the SIMT engine."""
    return algo == config.ALGO_TCGEN05_TF32 and k_pitch >= 32


class _StyledConvFn(Function):
    """a = [lrelu]( d * conv(s*x, c*W) [+ nw*noise] [+ bias] ) on NHWC-p buffers.

    Inputs: x [B,I,H,W]; s [B,I]; d [B,O] or None; weight [1,O,I,k,k]; noise [B|1,1,Ho,Wo] or None;
    noise_w [1] or None; bias [O] or None.  Non-tensor: wscale (c), upsample, fir [4,4], pad, act.
    """

    @staticmethod
    def forward(ctx, x, s, d, weight, noise, noise_w, bias, wscale, upsample, fir, pad, act, algo):
        require_cuda(x, 'ModulatedConv2d')
        b, cin, h, w = x.shape
        _, cout, cin_w, k, _ = weight.shape
        if cin_w != cin:
            raise RuntimeError(f'ModulatedConv2d: input has {cin} channels, weight expects {cin_w}')
        pin, pout = pitch_of(cin), pitch_of(cout)
        dev = x.device
        with torch.cuda.device(dev):
            st = stream_of(x)
            xb = as_nhwc_buf(x.detach())
            s_p = _pad_last(s.detach().float(), pin)
            d_p = _pad_last(d.detach().float(), pout) if d is not None else None
            bias_p = _pad_last(bias.detach().float(), pout) if bias is not None else None
            wk = cached_frozen(weight, ('wk', wscale), lambda: weight.detach()[0] * wscale)   # [O,I,k,k]
            tc = _use_tc(algo, pin)
            if tc:
                # tensor-pipe path: operands come straight from TMA, so modulation is a tensor pass
                w_fwd = cached_frozen(weight, ('fwd_tc', wscale, pin),
                                      lambda: _slabs_tc(wk.permute(2, 3, 0, 1), pin))   # [t][o][i], K-major
                x_in = torch.empty_like(xb)
                if xb.numel():
                    check(lib.cagc_modulate(st, xb.data_ptr(), s_p.data_ptr(), x_in.data_ptr(), b, h, w, pin), 'modulate')
                s_arg, falgo = None, config.ALGO_TCGEN05_TF32
            else:
                w_fwd = cached_frozen(weight, ('fwd_simt', wscale, pin, pout),
                                      lambda: _slabs(wk.permute(2, 3, 1, 0), pin, pout))  # [t][i][o]
                x_in, s_arg, falgo = xb, s_p.data_ptr(), config.ALGO_SIMT_FP32
            if noise is not None:
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
                if out.numel():
                    flops = 2.0 * b * h * w * cin * cout * k * k          # Util/Calculators.py convention x2
                    _timed(f'conv_same[algo{falgo}]', flops, 4.0 * b * h * w * (cin + cout),
                           lambda: check(lib.cagc_conv_same(st, x_in.data_ptr(), w_fwd.data_ptr(), s_arg,
                                                            ptr(d_p), ptr(noise), ptr(nw), ptr(bias_p), out.data_ptr(),
                                                            b, h, w, pin, pout, cout, k, nstride, int(act), falgo),
                                         'conv_same'))
            else:
                hu, wu = 2 * h + k - 2, 2 * w + k - 2
                ut = torch.empty((b, hu, wu, pout), device=dev, dtype=torch.float32)
                fir = fir.contiguous()
                kh, kw = fir.shape
                ho, wo = hu + pad[0] + pad[1] - kh + 1, wu + pad[0] + pad[1] - kw + 1
                out = torch.empty((b, ho, wo, pout), device=dev, dtype=torch.float32)
                if out.numel():
                    flops = 2.0 * b * h * w * cin * cout * k * k          # transposed conv counted at input res
                    _timed(f'conv_up[algo{falgo}]', flops, 4.0 * b * (h * w * cin + hu * wu * cout),
                           lambda: check(lib.cagc_conv_up(st, x_in.data_ptr(), w_fwd.data_ptr(), s_arg,
                                                          ut.data_ptr(), b, h, w, pin, pout, k, falgo), 'conv_up'))
                    _timed('fir_nhwc', 0.0, 4.0 * b * cout * (hu * wu + ho * wo),
                           lambda: check(lib.cagc_fir_nhwc(st, ut.data_ptr(), fir.data_ptr(), ptr(d_p), ptr(noise),
                                                           ptr(nw), ptr(bias_p), out.data_ptr(), b, hu, wu, pout, cout,
                                                           kh, kw, pad[0], pad[1], pad[0], pad[1], nstride, int(act)),
                                         'fir_nhwc'))
                del ut
        xm = x_in if tc else None      # modulated, TF32-rounded input: A operand of the tensor-pipe wgrad
        ctx.save_for_backward(xb, s_p, d_p, wk, noise, nw, bias_p, out, fir if upsample else None, xm)
        ctx.cfg = (b, cin, cout, h, w, k, pin, pout, upsample, pad, bool(act), nstride, algo, wscale,
                   d is not None, bias is not None)
        return nhwc_view(out, cout)

    @staticmethod
    def backward(ctx, ga):
        xb, s_p, d_p, wk, noise, nw, bias_p, out, fir, xm = ctx.saved_tensors
        (b, cin, cout, h, w, k, pin, pout, upsample, pad, act, nstride, algo, wscale, has_d, has_bias) = ctx.cfg
        if torch.is_grad_enabled() and ga.requires_grad:
            raise RuntimeError('double backward through the fused modulated convolution is not implemented; '
                               'use b200gan.config.second_order() (composite path) for the PPL regulariser')
        dev = xb.device
        ho, wo = out.shape[1], out.shape[2]
        need_x, need_s, need_d, need_w = ctx.needs_input_grad[0:4]
        need_nw, need_bias = ctx.needs_input_grad[5], ctx.needs_input_grad[6]
        with torch.cuda.device(dev):
            st = stream_of(xb)
            gu = torch.empty_like(out)
            chunks = lib.cagc_act_bwd_chunks(ho, wo)
            partial = torch.empty((b, chunks, 3, pout), device=dev, dtype=torch.float32)
            sb, sc, sh, sw = ga.stride()
            check(lib.cagc_act_bwd(st, ga.data_ptr(), sb, sc, sh, sw, out.data_ptr(), ptr(d_p), ptr(noise), ptr(nw),
                                   ptr(bias_p), gu.data_ptr(), partial.data_ptr(), b, ho, wo, pout, cout,
                                   nstride, int(act)), 'act_bwd')
            sums = partial.sum(1)                                   # [B,3,P]
            g_bias = sums[:, 0, :cout].sum(0) if (has_bias and need_bias) else None
            g_d = sums[:, 1, :cout].contiguous() if (has_d and need_d) else None
            g_nw = sums[:, 2, :cout].sum().reshape(1) if (noise is not None and need_nw) else None

            if upsample:
                # blur backward: correlation with the FIR kernel, pads of op/upfirdn2d.py:111-116
                kh, kw = fir.shape
                hu, wu = 2 * h + k - 2, 2 * w + k - 2
                gp0, gp1 = kh - pad[0] - 1, hu - ho + pad[0]
                g_t = torch.empty((b, hu, wu, pout), device=dev, dtype=torch.float32)
                firf = torch.flip(fir, [0, 1]).contiguous()
                check(lib.cagc_fir_nhwc(st, gu.data_ptr(), firf.data_ptr(), None, None, None, None, g_t.data_ptr(),
                                        b, ho, wo, pout, pout, kh, kw, gp0, gp1, gp0, gp1, 0, 0), 'fir_nhwc(bwd)')
                g_conv = g_t
            else:
                g_conv = gu

            g_x = g_s = g_w = None
            if need_x or need_s:
                gxt = torch.empty((b, h, w, pin), device=dev, dtype=torch.float32)
                if upsample and _use_tc(algo, pout):
                    w_d = _slabs_tc(wk.permute(2, 3, 1, 0), pout)                    # [t][i][o], K = o
                    _timed('conv_up_dgrad[algo1]', 2.0 * b * h * w * cin * cout * k * k,
                           4.0 * b * (h * w * cin + hu * wu * cout),
                           lambda: check(lib.cagc_conv_up_dgrad(st, g_conv.data_ptr(), w_d.data_ptr(), gxt.data_ptr(),
                                                                b, h, w, pout, pin, k, config.ALGO_TCGEN05_TF32),
                                         'conv_up_dgrad(tc)'))
                elif upsample:
                    w_d = _slabs(wk.permute(2, 3, 0, 1), pout, pin)                 # [t][o][i]
                    _timed('conv_up_dgrad[algo0]', 2.0 * b * h * w * cin * cout * k * k,
                           4.0 * b * (h * w * cin + hu * wu * cout),
                           lambda: check(lib.cagc_conv_up_dgrad(st, g_conv.data_ptr(), w_d.data_ptr(), gxt.data_ptr(),
                                                                b, h, w, pout, pin, k, config.ALGO_SIMT_FP32),
                                         'conv_up_dgrad'))
                elif _use_tc(algo, pout):
                    w_d = _slabs_tc(torch.flip(wk, [2, 3]).permute(2, 3, 1, 0), pout)   # [t'][i][o], K = o
                    _timed('conv_same[algo1]', 2.0 * b * h * w * cin * cout * k * k, 4.0 * b * h * w * (cin + cout),
                           lambda: check(lib.cagc_conv_same(st, g_conv.data_ptr(), w_d.data_ptr(), None, None, None,
                                                            None, None, gxt.data_ptr(), b, h, w, pout, pin, pin, k, 0,
                                                            0, config.ALGO_TCGEN05_TF32), 'conv_same(dgrad,tc)'))
                else:
                    w_d = _slabs(torch.flip(wk, [2, 3]).permute(2, 3, 0, 1), pout, pin)
                    _timed('conv_same[algo0]', 2.0 * b * h * w * cin * cout * k * k, 4.0 * b * h * w * (cin + cout),
                           lambda: check(lib.cagc_conv_same(st, g_conv.data_ptr(), w_d.data_ptr(), None, None, None,
                                                            None, None, gxt.data_ptr(), b, h, w, pout, pin, pin, k, 0,
                                                            0, config.ALGO_SIMT_FP32), 'conv_same(dgrad)'))
                mchunks = lib.cagc_act_bwd_chunks(h, w)
                mpartial = torch.empty((b, mchunks, pin), device=dev, dtype=torch.float32)
                check(lib.cagc_mod_bwd(st, gxt.data_ptr(), xb.data_ptr(), s_p.data_ptr(), mpartial.data_ptr(),
                                       b, h, w, pin), 'mod_bwd')
                if need_s:
                    g_s = mpartial.sum(1)[:, :cin]
                if need_x:
                    g_x = nhwc_view(gxt, cin)
            if need_w:
                # exact-fp32 mode (saliency): SIMT engine; tensor-pipe mode: tcgen05 TF32.  Both reduce the
            # split-K partials in a fixed order (deterministic)
                walgo = config.ALGO_TCGEN05_TF32 if (xm is not None and _use_tc(algo, min(pin, pout))
                                                     and pout <= 256) else config.ALGO_SIMT_FP32
                nsp = lib.cagc_conv_wgrad_splits(b, h, w, pin, pout, k, walgo)
                gw = torch.empty((k * k, pin, pout), device=dev, dtype=torch.float32)
                wpart = torch.empty((nsp, k * k, pin, pout), device=dev, dtype=torch.float32)
                a_in, a_sc = (xm, None) if walgo == config.ALGO_TCGEN05_TF32 else (xb, s_p.data_ptr())
                _timed(f'conv_wgrad[algo{walgo}]', 2.0 * b * h * w * cin * cout * k * k,
                       4.0 * b * h * w * (cin + cout),
                       lambda: check(lib.cagc_conv_wgrad(st, a_in.data_ptr(), a_sc, g_conv.data_ptr(),
                                                         gw.data_ptr(), wpart.data_ptr(), nsp, b, h, w, pin, pout, k,
                                                         1 if upsample else 0, walgo), 'conv_wgrad'))
                g_w = (gw[:, :cin, :cout].reshape(k, k, cin, cout).permute(3, 2, 0, 1) * wscale).unsqueeze(0)
        return (g_x, g_s, g_d, g_w, None, g_nw, g_bias, None, None, None, None, None, None)


def styled_conv(x, s, d, weight, noise, noise_w, bias, wscale, upsample=False, fir=None, pad=(0, 0), act=True):
    return _StyledConvFn.apply(x, s, d, weight, noise, noise_w, bias, wscale, upsample, fir, pad, act,
                               config.conv_algo())


class _ToRGBFn(Function):
    """rgb = conv1x1(s*x, c*W) + bias + Upsample(skip), output NCHW [B,3,H,W] (model.py:380-395)."""

    @staticmethod
    def forward(ctx, x, s, weight, bias, skip, wscale, fir, pad):
        require_cuda(x, 'ToRGB')
        b, cin, h, w = x.shape
        nout = weight.shape[1]
        pin = pitch_of(cin)
        dev = x.device
        with torch.cuda.device(dev):
            st = stream_of(x)
            xb = as_nhwc_buf(x.detach())
            s_p = _pad_last(s.detach().float(), pin)
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
        return out

    @staticmethod
    def backward(ctx, g):
        xb, s_p, w2, fir = ctx.saved_tensors
        b, cin, nout, h, w, pin, wscale, pad, has_bias, has_skip = ctx.cfg
        if torch.is_grad_enabled() and g.requires_grad:
            raise RuntimeError('double backward through the fused ToRGB is not implemented; '
                               'use b200gan.config.second_order()')
        dev = xb.device
        g = g.contiguous()
        with torch.cuda.device(dev):
            st = stream_of(xb)
            gx = torch.empty((b, h, w, pin), device=dev, dtype=torch.float32)
            chunks = lib.cagc_act_bwd_chunks(h, w)
            partial = torch.empty((b, chunks, nout, pin), device=dev, dtype=torch.float32)
            check(lib.cagc_torgb_bwd(st, g.data_ptr(), xb.data_ptr(), w2.data_ptr(), s_p.data_ptr(), gx.data_ptr(),
                                     partial.data_ptr(), b, h, w, pin, cin, nout, wscale), 'torgb_bwd')
            t = partial.sum(1)[:, :, :cin]                                   # [B,nout,I] = sum_p g*x
            s = s_p[:, :cin]
            g_w = g_s = g_bias = g_skip = None
            if ctx.needs_input_grad[2]:
                g_w = (wscale * torch.einsum('boi,bi->oi', t, s)).reshape(1, nout, cin, 1, 1)
            if ctx.needs_input_grad[1]:
                g_s = wscale * torch.einsum('boi,oi->bi', t, w2)
            if has_bias and ctx.needs_input_grad[3]:
                g_bias = g.sum(dim=(0, 2, 3)).reshape(1, nout, 1, 1)
            if has_skip and ctx.needs_input_grad[4]:
                from op.upfirdn2d import _launch
                fh, fw = fir.shape
                # backward of Upsample(up=2, pad): down=2 with the flipped kernel (op/upfirdn2d.py:111-116)
                gp0 = fh - pad[0] - 1
                gp1 = (h // 2) * 2 - h + pad[0] - 2 + 1
                g_skip = _launch(g, torch.flip(fir, [0, 1]).contiguous(), (1, 1), (2, 2), (gp0, gp1, gp0, gp1))
            g_x = nhwc_view(gx, cin) if ctx.needs_input_grad[0] else None
        return g_x, g_s, g_w, g_bias, g_skip, None, None, None


def to_rgb(x, s, weight, bias, skip, wscale, fir=None, pad=(0, 0)):
    return _ToRGBFn.apply(x, s, weight, bias, skip, wscale, fir, pad)


def demod_coefficients(s: torch.Tensor, weight: torch.Tensor, wscale: float, eps: float = 1e-8) -> torch.Tensor:
    """d[b,o] = rsqrt(sum_i s[b,i]^2 * Wsq[o,i] + eps), Wsq = sum_taps (c*W)^2   (model.py:251-253)."""
    wsq = cached_frozen(weight, ('wsq', wscale), lambda: (weight[0] * wscale).square().sum(dim=(2, 3)))   # [O,I]
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
