"""`upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))` -- public surface of op/upfirdn2d.py:145-156.

Forward, backward and double backward all run the same native entry point
(`cagc_upfirdn2d`): the gradient of upfirdn2d is upfirdn2d with the flipped
kernel, up and down swapped and the pads of op/upfirdn2d.py:111-116.
"""
import torch
from torch.autograd import Function

from b200gan._lib import lib, check, stream_of, require_cuda, fir_nhwc


def _is_nhwc(x):
    return x.ndim == 4 and x.shape[1] > 1 and x.is_contiguous(memory_format=torch.channels_last) \
        and not x.is_contiguous()


def _dense(x, like=None):
    """Dense storage for the native call: channels-last stays channels-last (no layout round trip
    around cuDNN's NHWC kernels in the discriminator), everything else becomes NCHW-contiguous."""
    if like is not None and _is_nhwc(like) and x.shape[1] % 4 == 0:
        return x.contiguous(memory_format=torch.channels_last)
    if _is_nhwc(x) and x.shape[1] % 4 == 0:
        return x
    return x.contiguous()


def _launch(x4, kernel, up, down, pad):
    """x4: [N, C, H, W] fp32 CUDA, dense NCHW or dense channels-last.  Returns [N, C, out_h, out_w]
    in the same memory format."""
    up_x, up_y = up
    down_x, down_y = down
    pad_x0, pad_x1, pad_y0, pad_y1 = pad
    n, c, in_h, in_w = x4.shape
    kh, kw = kernel.shape
    out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) // down_y + 1
    out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) // down_x + 1
    nhwc = _is_nhwc(x4)
    if nhwc and not (up == (1, 1) and down == (1, 1) and (kh, kw) == (4, 4) and c % 4 == 0):
        x4, nhwc = x4.contiguous(), False
    out = torch.empty((n, c, max(out_h, 0), max(out_w, 0)), device=x4.device, dtype=x4.dtype,
                      memory_format=torch.channels_last if nhwc else torch.contiguous_format)
    if out.numel() == 0:
        return out
    if nhwc:
        # channels-last storage is exactly the NHWC-p layout with pitch == C
        with torch.cuda.device(x4.device):
            fir_nhwc(stream_of(x4), x4.data_ptr(), kernel, None, None, None, None, out.data_ptr(), n, in_h, in_w, c, c,
                     (pad_x0, pad_x1, pad_y0, pad_y1), 0, 0, 'upfirdn2d(nhwc)')
        return out
    with torch.cuda.device(x4.device):
        check(lib.cagc_upfirdn2d(stream_of(x4), x4.data_ptr(), kernel.data_ptr(), out.data_ptr(),
                                 n * c, in_h, in_w, 1, kh, kw, up_x, up_y, down_x, down_y,
                                 pad_x0, pad_x1, pad_y0, pad_y1), 'upfirdn2d')
    return out


class _UpFirDn2dBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, kernel, kernel_flipped, up, down, pad, g_pad):
        ctx.save_for_backward(kernel)
        ctx.cfg = (up, down, pad)
        grad_input = _launch(_dense(grad_output), kernel_flipped, down, up, g_pad)
        return grad_input

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        up, down, pad = ctx.cfg
        # d(grad_input)/d(grad_output) is the forward operator itself (op/upfirdn2d.py:62-85)
        gradgrad_out = _UpFirDn2d.apply(gradgrad_input, kernel, up, down, pad)
        return gradgrad_out, None, None, None, None, None, None


class _UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kh, kw = kernel.shape
        _, _, in_h, in_w = input.shape
        out = _launch(_dense(input), kernel, up, down, pad)
        out_h, out_w = out.shape[2:]
        g_pad = (kw - pad_x0 - 1, in_w * up_x - out_w * down_x + pad_x0 - up_x + 1,
                 kh - pad_y0 - 1, in_h * up_y - out_h * down_y + pad_y0 - up_y + 1)
        # flip(kernel) is cached per (long-lived) FIR buffer: a fresh tensor per call would miss the identity-keyed
        # host-tap cache in the backward and cost one blocking device->host copy per Blur backward
        from b200gan.modconv import _flipped
        ctx.save_for_backward(kernel, _flipped(kernel))
        ctx.cfg = (up, down, pad, g_pad)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        kernel, kernel_flipped = ctx.saved_tensors
        up, down, pad, g_pad = ctx.cfg
        grad_input = _UpFirDn2dBackward.apply(grad_output, kernel, kernel_flipped, up, down, pad, g_pad)
        return grad_input, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    require_cuda(input, 'upfirdn2d')
    if input.ndim != 4 or kernel.ndim != 2:
        raise RuntimeError('upfirdn2d: expected input [N,C,H,W] and kernel [kh,kw]')
    kernel = kernel.to(device=input.device, dtype=torch.float32).contiguous()
    return _UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
