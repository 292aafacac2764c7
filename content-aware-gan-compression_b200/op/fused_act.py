"""`FusedLeakyReLU` / `fused_leaky_relu` -- public surface of op/fused_act.py:87-119.

y = leaky_relu(x + bias[channel], negative_slope) * scale, bias indexed along dim 1.
Works on NCHW-contiguous and on channels-last (dense NHWC) storage without a copy: the native
kernel only needs the (step_b, size_b) geometry of the bias index (op/fused_bias_act_kernel.cu:33).
Backward reduces the bias gradient in the same pass; double backward (R1 / path-length
regularisers, op/fused_act.py:47-53) is the same kernel again.
"""
import torch
from torch import nn
from torch.autograd import Function

from b200gan._lib import lib, check, stream_of, require_cuda, ptr


def _geometry(x):
    """Return (dense tensor to run on, step_b, size_b) for bias along dim 1."""
    if x.ndim < 2:
        raise RuntimeError('fused_leaky_relu: input needs at least 2 dims')
    if x.ndim == 4 and x.shape[1] > 1 and x.is_contiguous(memory_format=torch.channels_last) \
            and not x.is_contiguous():
        return x, 1, x.shape[1]
    x = x.contiguous()
    step = 1
    for s in x.shape[2:]:
        step *= s
    return x, step, x.shape[1]


def _bias_act(x, bias, refer, act, grad, alpha, scale, step_b, size_b):
    out = torch.empty_like(x)  # preserves the dense layout of x
    if x.numel() == 0:
        return out
    with torch.cuda.device(x.device):
        check(lib.cagc_fused_bias_act(stream_of(x), x.data_ptr(), ptr(bias), ptr(refer), out.data_ptr(),
                                      x.numel(), step_b, size_b, act, grad, alpha, scale), 'fused_bias_act')
    return out


class _FusedLeakyReLUBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, out, has_bias, negative_slope, scale, step_b, size_b):
        ctx.save_for_backward(out)
        ctx.cfg = (negative_slope, scale, step_b, size_b)
        if step_b == 1:
            g = grad_output.contiguous(memory_format=torch.channels_last) if out.ndim == 4 and not out.is_contiguous() \
                else grad_output.contiguous()
        else:
            g = grad_output.contiguous()
        rows_chunks = lib.cagc_bias_grad_rows_chunks(g.numel() // size_b, size_b) if (has_bias and step_b == 1) else 0
        if rows_chunks > 0 and g.numel() > 0:
            grad_input = torch.empty_like(g)
            partial = torch.empty((rows_chunks, size_b), device=g.device, dtype=g.dtype)
            with torch.cuda.device(g.device):
                check(lib.cagc_fused_bias_act_bwd_rows(stream_of(g), g.data_ptr(), out.data_ptr(), grad_input.data_ptr(),
                                                       partial.data_ptr(), g.numel() // size_b, size_b, negative_slope,
                                                       scale), 'fused_bias_act_bwd_rows')
            return grad_input, partial.sum(dim=0)
        chunks = lib.cagc_bias_grad_chunks(step_b) if has_bias else 0
        if chunks > 0 and g.numel() > 0 and (g.numel() // step_b) <= 65535:
            grad_input = torch.empty_like(g)
            outer = g.numel() // (step_b * size_b)
            partial = torch.empty((outer, size_b, chunks), device=g.device, dtype=g.dtype)
            with torch.cuda.device(g.device):
                check(lib.cagc_fused_bias_act_bwd(stream_of(g), g.data_ptr(), out.data_ptr(), grad_input.data_ptr(),
                                                  partial.data_ptr(), outer, size_b, step_b, negative_slope, scale),
                      'fused_bias_act_bwd')
            grad_bias = partial.sum(dim=(0, 2))
        else:
            grad_input = _bias_act(g, None, out, 3, 1, negative_slope, scale, step_b, size_b)
            if has_bias:
                dims = [0] + list(range(2, grad_input.ndim))
                grad_bias = grad_input.sum(dims).detach()
            else:
                grad_bias = grad_input.new_empty(0)
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        negative_slope, scale, step_b, size_b = ctx.cfg
        gg, _, _ = _match_layout(gradgrad_input, out)
        bias = gradgrad_bias.contiguous() if gradgrad_bias is not None and gradgrad_bias.numel() else None
        gradgrad_out = _bias_act(gg, bias, out, 3, 1, negative_slope, scale, step_b, size_b)
        return gradgrad_out, None, None, None, None, None, None


def _match_layout(t, like):
    if like.ndim == 4 and not like.is_contiguous():
        return t.contiguous(memory_format=torch.channels_last), None, None
    return t.contiguous(), None, None


class _FusedLeakyReLU(Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        x, step_b, size_b = _geometry(input)
        if bias is not None:
            if bias.numel() != size_b:
                raise RuntimeError(f'fused_leaky_relu: bias has {bias.numel()} elements, dim 1 has {size_b}')
            bias = bias.contiguous()
        out = _bias_act(x, bias, None, 3, 0, negative_slope, scale, step_b, size_b)
        ctx.save_for_backward(out)
        ctx.cfg = (bias is not None, negative_slope, scale, step_b, size_b)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        has_bias, negative_slope, scale, step_b, size_b = ctx.cfg
        grad_input, grad_bias = _FusedLeakyReLUBackward.apply(grad_output, out, has_bias, negative_slope, scale,
                                                             step_b, size_b)
        return grad_input, (grad_bias if has_bias else None), None, None


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    require_cuda(input, 'fused_leaky_relu')
    return _FusedLeakyReLU.apply(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel)) if bias else None
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)
