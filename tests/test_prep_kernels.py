"""GPU parity of the layer-bookkeeping kernels (csrc/prep.cu) against plain fp32/fp64 torch formulas of the
same quantities (model.py:248-257 and SURVEY.md App. B), called through the C ABI."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def env():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    from b200gan import _lib, modconv
    return _lib, modconv


def _close(a, b, tol=2e-6):
    a, b = a.double().cpu(), b.double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
    assert err < tol, err


@pytest.mark.parametrize('O,I,k,up,tc', [(7, 10, 3, False, False), (39, 77, 3, True, True), (154, 154, 3, False, True),
                                          (3, 39, 1, False, False), (20, 39, 3, False, True)])
def test_weight_prep(env, O, I, k, up, tc):
    _lib, mc = env
    torch.manual_seed(0)
    dev = 'cuda'
    w = torch.randn(1, O, I, k, k, device=dev)
    bias = torch.randn(O, device=dev)
    c = 1.0 / math.sqrt(I * k * k)
    pin, pout = mc.pitch_of(I), mc.pitch_of(O)
    tc_f, tc_d = tc and pin >= 32, tc and pout >= 32
    r = mc._weight_prep(w, bias, c, up, tc_f, tc_d, pin, pout, True)
    wk = w[0] * c

    def tf32(t):
        out = torch.empty_like(t)
        _lib.check(_lib.lib.cagc_modulate(None, t.data_ptr(), None, out.data_ptr(), 1, 1, t.numel() // 4, 4))
        return out

    def pad2(t, rows, cols):
        return torch.nn.functional.pad(t, (0, cols - t.shape[-1], 0, rows - t.shape[-2])).contiguous()

    if tc_f:
        ref_f = tf32(pad2(wk.permute(2, 3, 0, 1).reshape(k * k, O, I), mc._r16(pout), pin))
    else:
        ref_f = pad2(wk.permute(2, 3, 1, 0).reshape(k * k, I, O), pin, pout)
    wd = wk if up else torch.flip(wk, [2, 3])
    if tc_d:
        ref_d = tf32(pad2(wd.permute(2, 3, 1, 0).reshape(k * k, I, O), mc._r16(pin), pout))
    else:
        ref_d = pad2(wd.permute(2, 3, 0, 1).reshape(k * k, O, I), pout, pin)
    assert torch.equal(r.w_fwd, ref_f)
    assert torch.equal(r.w_dgrad, ref_d)
    wsq = pad2(wk.square().sum(dim=(2, 3)), pout, pin)
    _close(r.wsq_oi, wsq)
    _close(r.wsq_io, wsq.t())
    assert torch.equal(r.bias_p, torch.nn.functional.pad(bias, (0, pout - O)))


def test_style_affine_and_demod(env):
    _lib, mc = env
    torch.manual_seed(1)
    dev = 'cuda'
    B, L, D = 5, 6, 96

    class Mod:
        def __init__(self, i, bias=True):
            self.weight = torch.randn(i, D, device=dev, requires_grad=True)
            self.bias = torch.randn(i, device=dev, requires_grad=True) if bias else None
            self.scale = 1 / math.sqrt(D)
            self.lr_mul = 1.0

    mods = [Mod(39), Mod(154), Mod(8, bias=False), Mod(77)]
    rows = [0, 3, 3, 5]
    latent = torch.randn(B, L + 2, D, device=dev)[:, 1:L + 1].requires_grad_(True)   # strided view
    outs = mc.style_affine(latent, mods, rows)
    cot = [torch.randn_like(o) for o in outs]
    loss = sum((o * c_).sum() for o, c_ in zip(outs, cot))
    params = [latent] + [m.weight for m in mods] + [m.bias for m in mods if m.bias is not None]
    got = torch.autograd.grad(loss, params)
    ref_outs = []
    for m, r_ in zip(mods, rows):
        s = latent[:, r_].double() @ (m.weight.double() * m.scale).t()
        if m.bias is not None:
            s = s + m.bias.double() * m.lr_mul
        ref_outs.append(s)
    for o, r_, m in zip(outs, ref_outs, mods):
        i = m.weight.shape[0]
        _close(o[:, :i], r_, 1e-5)
        assert float(o[:, i:].abs().sum()) == 0.0
    ref_loss = sum((r_ * c_[:, :r_.shape[1]].double()).sum() for r_, c_ in zip(ref_outs, cot))
    ref = torch.autograd.grad(ref_loss, params)
    for a, b in zip(got, ref):
        _close(a, b, 1e-5)

    # demodulation coefficients
    s_p = outs[1].detach()
    wsq = torch.rand(154 + 6, 160, device=dev)      # [pin][pout]
    wsq[154:] = 0
    d = torch.empty(B, 160, device=dev)
    _lib.check(_lib.lib.cagc_demod(None, s_p.data_ptr(), wsq.data_ptr(), d.data_ptr(), B, 160, 154, 160, 1e-8))
    ref_d = torch.rsqrt(s_p.double().square() @ wsq.double() + 1e-8)
    _close(d[:, :154], ref_d[:, :154], 1e-5)
    assert float(d[:, 154:].abs().sum()) == 0.0


def test_backward_finalizers(env):
    _lib, mc = env
    lib, check = _lib.lib, _lib.check
    torch.manual_seed(2)
    dev = 'cuda'
    B, chunks, O, I, k = 6, 5, 39, 77, 3
    pin, pout = mc.pitch_of(I), mc.pitch_of(O)
    partial = torch.randn(B, chunks, 3, pout, device=dev)
    d = torch.rand(B, pout, device=dev) + 0.5
    d[:, O:] = 0
    g_bias = torch.empty(pout, device=dev)
    gq = torch.empty(B, pout, device=dev)
    nblk = lib.cagc_act_bwd_finalize_blocks(pout)
    nwp = torch.empty(nblk, device=dev)
    check(lib.cagc_act_bwd_finalize(None, partial.data_ptr(), d.data_ptr(), g_bias.data_ptr(), gq.data_ptr(),
                                    nwp.data_ptr(), B, chunks, pout))
    sums = partial.double().sum(1)
    _close(g_bias, sums[:, 0].sum(0), 1e-5)
    _close(gq, -0.5 * d.double() ** 3 * sums[:, 1], 1e-5)
    _close(nwp.sum(), sums[:, 2].sum(), 1e-5)

    s = torch.randn(B, pin, device=dev)
    s[:, I:] = 0
    wsq = torch.rand(pout, pin, device=dev)
    mchunks = 4
    mpart = torch.randn(B, mchunks, pin, device=dev)
    g_s = torch.empty(B, pin, device=dev)
    check(lib.cagc_style_grad_finalize(None, mpart.data_ptr(), gq.data_ptr(), s.data_ptr(), wsq.data_ptr(),
                                       g_s.data_ptr(), B, mchunks, pin, O, pout))
    ref = mpart.double().sum(1) + 2 * s.double() * (gq.double()[:, :O] @ wsq.double()[:O])
    _close(g_s, ref, 1e-5)
    check(lib.cagc_style_grad_finalize(None, mpart.data_ptr(), None, s.data_ptr(), None, g_s.data_ptr(), B, mchunks,
                                       pin, O, pout))
    _close(g_s, mpart.double().sum(1), 1e-5)

    nsp = 3
    wpart = torch.randn(nsp, k * k, pin, pout, device=dev)
    w = torch.randn(O, I, k, k, device=dev)
    c = 0.37
    out = torch.empty(1, O, I, k, k, device=dev)
    check(lib.cagc_wgrad_finalize(None, wpart.data_ptr(), nsp, w.data_ptr(), c, gq.data_ptr(), s.data_ptr(), B, O, I, k,
                                  pin, pout, out.data_ptr()))
    gw = wpart.double().sum(0)[:, :I, :O].reshape(k, k, I, O).permute(3, 2, 0, 1) * c
    dem = 2 * c * c * w.double() * (gq.double()[:, :O].t() @ s.double()[:, :I].square())[:, :, None, None]
    _close(out[0], gw + dem, 1e-5)
    check(lib.cagc_wgrad_finalize(None, wpart.data_ptr(), nsp, w.data_ptr(), c, None, None, B, O, I, k, pin, pout,
                                  out.data_ptr()))
    _close(out[0], gw, 1e-5)

    nout = 3
    tpart = torch.randn(B, chunks, nout, pin, device=dev)
    w2 = torch.randn(nout, I, device=dev)
    g_w = torch.empty(B, nout, I, device=dev)
    g_s2 = torch.empty(B, pin, device=dev)
    check(lib.cagc_torgb_bwd_finalize(None, tpart.data_ptr(), s.data_ptr(), w2.data_ptr(), c, g_w.data_ptr(),
                                      g_s2.data_ptr(), B, chunks, I, pin, nout))
    t = tpart.double().sum(1)[:, :, :I]
    _close(g_w.sum(0), c * torch.einsum('boi,bi->oi', t, s.double()[:, :I]), 1e-5)
    _close(g_s2[:, :I], c * torch.einsum('boi,oi->bi', t, w2.double()), 1e-5)
    assert float(g_s2[:, I:].abs().sum()) == 0.0


@pytest.mark.parametrize('act', [False, True])
def test_equal_linear_fused(env, act):
    _lib, mc = env
    torch.manual_seed(3)
    dev = 'cuda'
    x = torch.randn(7, 48, device=dev, requires_grad=True)
    w = torch.randn(33, 48, device=dev, requires_grad=True)
    b = torch.randn(33, device=dev, requires_grad=True)
    scale, lr = 0.01 / math.sqrt(48), 0.01
    y = mc.equal_linear(x, w, b, scale, lr, act)
    cot = torch.randn_like(y)
    got = torch.autograd.grad((y * cot).sum(), [x, w, b])
    xd, wd, bd = x.double(), w.double(), b.double()
    r = xd @ (wd * scale).t() + bd * lr
    if act:
        r = torch.nn.functional.leaky_relu(r, 0.2) * math.sqrt(2)
    ref = torch.autograd.grad((r * cot.double()).sum(), [x, w, b])
    _close(y, r, 1e-5)
    for a, b_ in zip(got, ref):
        _close(a, b_, 1e-5)
    with pytest.raises(RuntimeError):
        x2 = torch.randn(3, 48, device=dev, requires_grad=True)
        g, = torch.autograd.grad(mc.equal_linear(x2, w, b, scale, lr, act).sum(), x2, create_graph=True)
        g.sum().backward()
