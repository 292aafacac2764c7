"""Discriminator path on the package's own engines (b200gan/dconv.py, csrc/disc.cu) against the fp64 CPU oracle:
every new entry point alone, then whole ResBlocks / the whole frozen Discriminator on both convolution engines at
BASELINE's 256px widths (128/256/512 channels)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synth

pytestmark = pytest.mark.gpu


def relmax(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _nhwc(t):
    """[B,C,H,W] fp64 cpu -> NHWC fp32 cuda buffer [B,H,W,C]."""
    return t.permute(0, 2, 3, 1).contiguous().float().cuda()


def _nchw(buf):
    return buf.permute(0, 3, 1, 2)


@pytest.mark.parametrize('up,down,pad,h,w,c', [(1, 2, (1, 1), 16, 16, 8), (1, 2, (1, 1), 34, 30, 24), (2, 1, (2, 1), 8, 8, 8),
                                               (2, 1, (2, 1), 17, 15, 40), (1, 2, (2, 2), 20, 20, 16)])
def test_fir_resample_nhwc(up, down, pad, h, w, c):
    import ctypes as C
    from b200gan._lib import lib, check
    from oracle import stylegan2_oracle as O
    g = torch.Generator().manual_seed(h * 100 + w + up)
    x = torch.randn(2, c, h, w, generator=g, dtype=torch.float64)
    k = O.fir_kernel_2d([1, 3, 3, 1]) * (up * up)
    ref = O.upfirdn2d(x, k, up=up, down=down, pad=pad)
    xb = _nhwc(x)
    out = torch.empty((2, ref.shape[2], ref.shape[3], c), device='cuda')
    taps = (C.c_float * 16)(*k.reshape(-1).tolist())
    check(lib.cagc_fir_resample_nhwc(torch.cuda.current_stream().cuda_stream, xb.data_ptr(), taps, out.data_ptr(), 2, h, w, c,
                                     up, down, pad[0], pad[1]))
    assert relmax(_nchw(out), ref) <= 2e-6


@pytest.mark.parametrize('algo', [0, 1])
@pytest.mark.parametrize('cin,cout,k,mode,h', [(32, 48, 3, 0, 12), (64, 32, 1, 0, 9), (40, 64, 3, 1, 17), (128, 256, 3, 1, 33),
                                               (520, 512, 3, 0, 4), (128, 128, 3, 0, 64), (128, 128, 3, 0, 128),
                                               (96, 160, 3, 0, 256), (256, 256, 3, 0, 128)])
def test_conv2d_modes_vs_torch(algo, cin, cout, k, mode, h):
    """cagc_conv2d (both engines) against F.conv2d in fp64: same-size and stride-2 convolutions, bias, activation
    gain, residual add after the activation."""
    from b200gan._lib import lib, check
    from b200gan.modconv import _weight_prep
    g = torch.Generator().manual_seed(cin + cout + k + h)
    b = 2
    x = torch.randn(b, cin, h, h, generator=g, dtype=torch.float64)
    wt = torch.randn(cout, cin, k, k, generator=g, dtype=torch.float64)
    bias = torch.randn(cout, generator=g, dtype=torch.float64)
    scale = 1.0 / math.sqrt(cin * k * k)
    if mode == 0:
        u = F.conv2d(x, wt * scale, padding=k // 2)
    else:
        u = F.conv2d(x, wt * scale, stride=2)
    res = torch.randn(u.shape, generator=g, dtype=torch.float64)
    ref = F.leaky_relu(u + bias.view(1, -1, 1, 1), 0.2) * 1.25 + res
    tc = bool(algo)
    prep = _weight_prep(wt.float().cuda().reshape(1, cout, cin, k, k), bias.float().cuda(), scale, mode == 1, tc, tc, cin, cout,
                        False)
    out = torch.empty((b, u.shape[2], u.shape[3], cout), device='cuda')
    resb, xb = _nhwc(res), _nhwc(x)
    check(lib.cagc_conv2d(torch.cuda.current_stream().cuda_stream, xb.data_ptr(), prep.w_fwd.data_ptr(),
                          prep.bias_p.data_ptr(), resb.data_ptr(), out.data_ptr(), b, h, h, cin, cout, cout, k, mode, 1, 1.25,
                          algo))
    e = relmax(_nchw(out), ref)
    assert e <= (3e-3 if tc else 2e-5), e


def test_from_rgb_and_act_mask():
    from b200gan._lib import lib, check
    g = torch.Generator().manual_seed(3)
    b, h, w, cout = 2, 19, 23, 128
    img = torch.randn(b, 3, h, w, generator=g, dtype=torch.float64)
    wt = torch.randn(cout, 3, generator=g, dtype=torch.float64)
    bias = torch.randn(cout, generator=g, dtype=torch.float64)
    scale = 1 / math.sqrt(3)
    imr = img.clone().requires_grad_(True)
    ref = F.leaky_relu(torch.einsum('oc,bchw->bohw', wt * scale, imr) + bias.view(1, -1, 1, 1), 0.2) * math.sqrt(2)
    cot = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    gref, = torch.autograd.grad(ref, imr, cot)
    st = torch.cuda.current_stream().cuda_stream
    wc, bc, cotb = wt.float().cuda(), bias.float().cuda(), _nhwc(cot)      # keep the device buffers alive
    for layout in ('nchw', 'nhwc'):
        ic = img.float().cuda()
        if layout == 'nhwc':
            ic = ic.contiguous(memory_format=torch.channels_last)
        y = torch.empty((b, h, w, cout), device='cuda')
        check(lib.cagc_from_rgb_fwd(st, ic.data_ptr(), *ic.stride(), wc.data_ptr(), bc.data_ptr(),
                                    y.data_ptr(), b, h, w, 3, cout, cout, scale, 1, math.sqrt(2)))
        assert relmax(_nchw(y), ref) <= 2e-6
    gimg = torch.empty((b, 3, h, w), device='cuda')
    check(lib.cagc_from_rgb_bwd(st, cotb.data_ptr(), y.data_ptr(), wc.data_ptr(), gimg.data_ptr(), b, h, w, 3,
                                cout, cout, scale, 1, math.sqrt(2)))
    assert relmax(gimg, gref) <= 5e-6
    gz = torch.empty_like(y)
    check(lib.cagc_act_mask_nhwc(st, cotb.data_ptr(), y.data_ptr(), gz.data_ptr(), gz.numel(), 0.7))
    expect = cot * 0.7 * torch.where(ref > 0, 1.0, 0.2)
    assert relmax(_nchw(gz), expect) <= 1e-6


def _oracle_resblock(O, sd, prefix, x):
    y = O._conv_layer(x, sd, f'{prefix}.conv1', 3)
    y = O._conv_layer(y, sd, f'{prefix}.conv2', 3, downsample=True)
    s = O._conv_layer(x, sd, f'{prefix}.skip', 1, downsample=True, activate=False, bias=False)
    return (y + s) / math.sqrt(2.0)


@pytest.mark.parametrize('cin,cout,h,b', [(32, 48, 16, 2), (128, 256, 64, 2), (512, 512, 8, 3)])
def test_resblock_vs_oracle_both_engines(cin, cout, h, b):
    """model.ResBlock with frozen parameters = one fused autograd node; forward and input gradient against the
    oracle's composition of model.py:670-737.  exact-fp32 engine <= 2e-5 / 1e-4, TF32 engine in L2 <= 3e-3 (forward) / 3e-2 (gradient)."""
    import model
    from b200gan import config
    from oracle import stylegan2_oracle as O
    blk = model.ResBlock(cin, cout)
    synth.load_synth(blk, cin + cout)
    sd = {f'blk.{k}': v.double() for k, v in blk.state_dict().items()}
    blk = blk.cuda()
    for p in blk.parameters():
        p.requires_grad_(False)
    g = torch.Generator().manual_seed(h)
    x = torch.randn(b, cin, h, h, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    ref = _oracle_resblock(O, sd, 'blk', xr)
    cot = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    gref, = torch.autograd.grad(ref, xr, cot)
    for algo in (config.ALGO_SIMT_FP32, config.ALGO_TCGEN05_TF32):
        xc = x.float().cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        with config.use_algo(algo):
            y = blk(xc)
            assert y.grad_fn is not None and 'ResBlockFn' in type(y.grad_fn).__name__, 'fused block not taken'
            gx, = torch.autograd.grad(y, xc, cot.float().cuda())
        if algo == config.ALGO_SIMT_FP32:
            assert relmax(y, ref) <= 2e-5 and relmax(gx, gref) <= 1e-4, (relmax(y, ref), relmax(gx, gref))
        else:
            # TF32 forward errors flip the sign of ~1e-3 of the near-zero pre-activations: an L2 effect of ~1e-2 on the gradient
            assert rel_l2(y, ref) <= 3e-3 and rel_l2(gx, gref) <= 3e-2, (rel_l2(y, ref), rel_l2(gx, gref))


def test_discriminator_256_frozen_vs_oracle():
    """The whole frozen 256px Discriminator (configs[1]'s D: 128 -> 256 -> 512 channels) on both engines."""
    import model
    from b200gan import config
    from oracle import stylegan2_oracle as O
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    d = model.Discriminator(256)
    synth.load_synth(d, 77)
    sd = {k: v.double() for k, v in d.state_dict().items()}
    d = d.cuda()
    for p in d.parameters():
        p.requires_grad_(False)
    g = torch.Generator().manual_seed(9)
    img = torch.randn(2, 3, 256, 256, generator=g, dtype=torch.float64)
    ir = img.clone().requires_grad_(True)
    ref = O.discriminator_forward(sd, ir, 256)
    gref, = torch.autograd.grad(F.softplus(-ref).mean(), ir)
    n0 = None
    for algo, tol_o, tol_g in ((config.ALGO_SIMT_FP32, 1e-4, 5e-3), (config.ALGO_TCGEN05_TF32, 2e-2, 5e-2)):
        ic = img.float().cuda().requires_grad_(True)
        with config.use_algo(algo):
            out = d(ic)
            gi, = torch.autograd.grad(F.softplus(-out).mean(), ic)
        assert relmax(out, ref) <= tol_o, (algo, relmax(out, ref))
        assert rel_l2(gi, gref) <= tol_g, (algo, rel_l2(gi, gref))
