"""Pin the CPU oracle (oracle/stylegan2_oracle.py) against outputs of the reference itself.

Fixtures: tests/golden/*.npz, produced by tests/golden/make_golden.py from the
unmodified reference in fp64.  Tolerance: 1e-10 relative-to-max (fp64, different
but equivalent summation order).
"""
import os

import numpy as np
import pytest
import torch

from oracle import stylegan2_oracle as O

TOL = 1e-10


def close(a, b, tol=TOL):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    denom = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / denom
    assert err <= tol, err


def T(a):
    return torch.from_numpy(np.asarray(a)).double()


def test_upfirdn2d_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'upfirdn2d.npz'))
    for i in range(int(g['n_cases'])):
        up, down, p0, p1 = [int(v) for v in g[f'c{i}_cfg']]
        x = T(g[f'c{i}_x']).requires_grad_(True)
        k = T(g[f'c{i}_k'])
        y = O.upfirdn2d(x, k, up, down, (p0, p1))
        close(y.detach(), g[f'c{i}_y'])
        gx, = torch.autograd.grad(y, x, T(g[f'c{i}_gy']))
        close(gx, g[f'c{i}_gx'])
        if x.numel() <= 400:
            close(O.upfirdn2d_numpy(g[f'c{i}_x'], g[f'c{i}_k'], up, down, (p0, p1)), g[f'c{i}_y'])


def test_upfirdn2d_backward_identity(golden_dir):
    """grad = upfirdn2d(gy, flip(k), up<->down, g_pad)  (op/upfirdn2d.py:111-116,19-44)."""
    g = np.load(os.path.join(golden_dir, 'upfirdn2d.npz'))
    for i in range(int(g['n_cases'])):
        up, down, p0, p1 = [int(v) for v in g[f'c{i}_cfg']]
        k = T(g[f'c{i}_k'])
        if k.shape[0] != k.shape[1]:
            continue
        x = g[f'c{i}_x']
        in_h, in_w = x.shape[2:]
        kh, kw = k.shape
        out_h, out_w = O.upfirdn2d_shape(in_h, in_w, kh, kw, up, down, p0, p1)
        if in_h != in_w:
            continue
        gp0 = kh - p0 - 1
        gp1 = in_h * up - out_h * down + p0 - up + 1
        gx = O.upfirdn2d(T(g[f'c{i}_gy']), torch.flip(k, [0, 1]), down, up, (gp0, gp1))
        close(gx, g[f'c{i}_gx'])


def test_fused_act_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'fused_act.npz'))
    for i in range(int(g['n_cases'])):
        x = T(g[f'c{i}_x']).requires_grad_(True)
        b = T(g[f'c{i}_b']).requires_grad_(True)
        y = O.fused_leaky_relu(x, b)
        close(y.detach(), g[f'c{i}_y'])
        gx, gb = torch.autograd.grad(y, [x, b], T(g[f'c{i}_gy']))
        close(gx, g[f'c{i}_gx'])
        close(gb, g[f'c{i}_gb'])
    close(O.fused_leaky_relu(T(g['nobias_x'])), g['nobias_y'])


@pytest.mark.parametrize('tag,up', [('same', False), ('up', True), ('same39', False)])
def test_styled_conv_golden(golden_dir, tag, up):
    g = np.load(os.path.join(golden_dir, 'layers.npz'))
    p = {'L.' + k[len(tag) + 4:]: T(g[k]).requires_grad_(True) for k in g.files if k.startswith(tag + '.sd.')}
    x = T(g[f'{tag}.x']).requires_grad_(True)
    w = T(g[f'{tag}.w']).requires_grad_(True)
    y = O.styled_conv(x, w, p, 'L', T(g[f'{tag}.noise']), upsample=up)
    close(y.detach(), g[f'{tag}.y'])
    names = sorted(k[len(tag) + 6:] for k in g.files if k.startswith(tag + '.grad.'))
    grads = torch.autograd.grad(y, [x, w] + [p['L.' + n] for n in names], T(g[f'{tag}.gy']))
    close(grads[0], g[f'{tag}.gx'])
    close(grads[1], g[f'{tag}.gw'])
    for n, gr in zip(names, grads[2:]):
        close(gr, g[f'{tag}.grad.{n}'])


def test_to_rgb_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'layers.npz'))
    p = {'R.' + k[len('rgb.sd.'):]: T(g[k]).requires_grad_(True) for k in g.files if k.startswith('rgb.sd.')}
    x = T(g['rgb.x']).requires_grad_(True)
    w = T(g['rgb.w']).requires_grad_(True)
    skip = T(g['rgb.skip']).requires_grad_(True)
    y = O.to_rgb(x, w, p, 'R', skip)
    close(y.detach(), g['rgb.y'])
    names = sorted(k[len('rgb.grad.'):] for k in g.files if k.startswith('rgb.grad.'))
    grads = torch.autograd.grad(y, [x, w, skip] + [p['R.' + n] for n in names], T(g['rgb.gy']))
    close(grads[0], g['rgb.gx'])
    close(grads[1], g['rgb.gw'])
    close(grads[2], g['rgb.gskip'])
    for n, gr in zip(names, grads[3:]):
        close(gr, g[f'rgb.grad.{n}'])


def _tiny(golden_dir):
    g = np.load(os.path.join(golden_dir, 'generator_tiny.npz'))
    sd = {k[3:]: T(g[k]) for k in g.files if k.startswith('sd.')}
    noise = [T(g[f'noise{i}']) for i in range(7)]
    return g, sd, noise


def test_generator_forward_golden(golden_dir):
    g, sd, noise = _tiny(golden_dir)
    z1, z2 = T(g['z1']), T(g['z2'])
    close(O.generator_forward(sd, 32, [z1], noise), g['img_single'])
    rgbs = O.generator_forward(sd, 32, [z1, z2], noise, inject_index=int(g['inject_index']), return_rgb_list=True)
    for i, r in enumerate(rgbs):
        close(r, g[f'rgb{i}'])
    stored = [sd[f'noises.noise_{i}'] for i in range(7)]
    close(O.generator_forward(sd, 32, [z1], stored), g['img_fixed_noise'])
    close(O.generator_forward(sd, 32, [z1], noise, truncation=0.7, truncation_latent=T(g['mean_w'])), g['img_trunc'])
    close(O.mapping_network(z1, sd, 2), g['w_latent'])
    close(O.generator_forward(sd, 32, [T(g['w_latent'])], noise, input_is_latent=True), g['img_from_w'])


def test_generator_grads_golden(golden_dir):
    g, sd, noise = _tiny(golden_dir)
    names = sorted(k[5:] for k in g.files if k.startswith('grad.'))
    p = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    rgbs = O.generator_forward(p, 32, [T(g['z1']), T(g['z2'])], noise, inject_index=int(g['inject_index']),
                               return_rgb_list=True)
    loss = (rgbs[-1] * T(g['cot'])).sum() + sum((r * r).mean() for r in rgbs[:-1])
    grads = torch.autograd.grad(loss, [p[n] for n in names])
    for n, gr in zip(names, grads):
        close(gr, g[f'grad.{n}'], 1e-9)


def _sp_fn(g):
    sel = torch.from_numpy(g['sal_mask'] & g['sal_hit'])
    val = torch.from_numpy(g['sal_pm1'].astype(np.float64))

    def fn(img):
        return torch.where(sel.view(1, 1, *sel.shape).expand_as(img), val.view(1, 1, *val.shape).expand_as(img), img)
    return fn


def test_saliency_and_prune_mask_golden(golden_dir):
    g, sd, noise = _tiny(golden_dir)
    scores = O.saliency_scores(sd, 32, T(g['z1']), noise, _sp_fn(g))
    assert len(scores) == 8
    for i, s in enumerate(scores):
        close(s, g[f'score{i}'], 1e-9)
    masks = O.prune_mask_from_scores(scores, 0.5)
    for i, m in enumerate(masks):
        assert np.array_equal(m, g[f'prune_mask{i}'])


def test_pruned_generator_golden(golden_dir):
    g, _, noise = _tiny(golden_dir)
    psd = {k[len('pruned_sd.'):]: T(g[k]) for k in g.files if k.startswith('pruned_sd.')}
    close(O.generator_forward(psd, 32, [T(g['z1'])], noise), g['img_pruned'])


def test_path_length_golden(golden_dir):
    g, sd, noise = _tiny(golden_dir)
    z1 = T(g['z1'])
    img, latent = O.generator_forward(sd, 32, [z1], noise, return_latent=True)
    # the reference differentiates w.r.t. the assembled latent (model.py:663); rebuild it as a leaf
    latent = latent.detach().requires_grad_(True)
    img = O.generator_forward(sd, 32, [latent], noise, input_is_latent=True)
    pl = O.path_lengths(img, latent, T(g['pl_noise']))
    close(pl.detach(), g['path_lengths'], 1e-9)


# ------------------------------------------------------------------ discriminator, KD-like step, RNG order
# (weights / inputs are re-drawn from seeds by tests/golden/synth.py, exactly as make_golden.py drew them for
#  the reference; only the reference's outputs live in the fixtures)
def _kd_tiny_inputs(g):
    import model          # the drop-in module tree supplies key order and shapes (no compute: CPU construction only)
    import synth
    from synth import KD_TINY as c
    disc_t = model.Discriminator(c['size']).state_dict()
    stu_t = model.Generator(c['size'], c['style_dim'], c['n_mlp'], generator_net_shape=c['student']).state_dict()
    tea_t = model.Generator(c['size'], c['style_dim'], c['n_mlp'], generator_net_shape=c['teacher']).state_dict()
    assert list(disc_t.keys()) == list(g['d_keys'])
    assert list(stu_t.keys()) == list(g['student_keys'])

    def sd(template, seed):
        out = {k: v.double() for k, v in template.items()}
        out.update({k: T(v) for k, v in synth.synth_state(template, seed).items()})
        return out
    dp, sp, tp = sd(disc_t, c['seed_disc']), sd(stu_t, c['seed_student']), sd(tea_t, c['seed_teacher'])
    n_layers = sum(1 for k in stu_t if k.startswith('noises.'))
    res = [stu_t[f'noises.noise_{i}'].shape[2:] for i in range(n_layers)]
    z, noise2, rs = synth.latents_and_noise(c['seed_inputs'], c['batch'], c['style_dim'],
                                            [tuple(r) for r in res] * 2, n_latents=2)
    z = [T(a) for a in z]
    s_noise, t_noise = [T(a) for a in noise2[:n_layers]], [T(a) for a in noise2[n_layers:]]
    return c, dp, sp, tp, z, s_noise, t_noise, rs


def test_discriminator_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'kd_tiny.npz'))
    c, dp, *_ = _kd_tiny_inputs(g)
    close(T(g['d_x']).numpy(), g['d_x'])
    x = T(g['d_x']).requires_grad_(True)
    pred = O.discriminator_forward(dp, x, c['size'])
    close(pred.detach(), g['d_pred'])
    gx, = torch.autograd.grad(pred, x, T(g['d_cot']))
    close(gx, g['d_gx'], 1e-9)
    close(O.discriminator_forward(dp, x[:2].detach(), c['size']), g['d_pred_b2'])


@pytest.mark.parametrize('mode', ['Output_Only', 'Intermediate'])
def test_kd_step_golden(golden_dir, mode):
    import synth
    g = np.load(os.path.join(golden_dir, 'kd_tiny.npz'))
    c, dp, sp, tp, z, s_noise, t_noise, _ = _kd_tiny_inputs(g)
    mask = T(synth.ellipse_mask(c['size']).astype(np.float64)).view(1, 1, c['size'], c['size'])
    loss, grads = O.kd_step(sp, tp, dp, c['size'], z, s_noise, t_noise, c['inject'], mask, kd_mode=mode)
    close(loss, g[f'{mode}.g_loss'] + g[f'{mode}.kd'])
    for n in g['param_names']:
        ref = g[f'{mode}.grad.{n}']
        got = grads[n].numpy() if n in grads else np.zeros_like(ref)
        if np.abs(ref).max() == 0:
            assert np.abs(got).max() == 0, n
        else:
            close(got, ref, 1e-9)


def test_rng_order_golden(golden_dir):
    """The noise list stored next to the image is the replay of one normal_() per layer in execution order
    (make_golden.gen_rng_order asserts the replay is bit-identical to the reference's internal draws)."""
    import model
    import synth
    from synth import TINY
    g = np.load(os.path.join(golden_dir, 'rng_order.npz'))
    tmpl = model.Generator(TINY['size'], TINY['style_dim'], TINY['n_mlp'], generator_net_shape=TINY['net_shape']).state_dict()
    sd = {k: v.double() for k, v in tmpl.items()}
    sd.update({k: T(v) for k, v in synth.synth_state(tmpl, int(g['seed_weights'])).items()})
    noise = [T(g[f'noise{i}']) for i in range(7)]
    close(O.generator_forward(sd, TINY['size'], [T(g['z'])], noise), g['img'])


# ------------------------------------------------------------------ rest of the KD loss: LPIPS-VGG16, content-mask glue
def test_lpips_oracle_matches_reference_lpips(golden_dir):
    """oracle.lpips_oracle.lpips_vgg against the reference's own lpips.PerceptualLoss (net-lin, vgg) in fp64."""
    import sys
    sys.path.insert(0, golden_dir)
    import synth
    from oracle import lpips_oracle as L
    g = np.load(os.path.join(golden_dir, 'lpips_tiny.npz'))
    c = synth.LPIPS_TINY
    cw, cb = synth.vgg16_weights(c['seed_weights'])
    cw, cb = [T(a) for a in cw], [T(a) for a in cb]
    lins = [T(g[f'lin{k}']) for k in range(5)]
    pred_np, target_np = synth.lpips_images(c['seed_inputs'], c['batch'], c['size'])
    pred = T(pred_np).requires_grad_(True)
    val = L.lpips_vgg(pred, T(target_np), cw, cb, lins)
    close(val.detach(), g['val'])
    gp, = torch.autograd.grad(val, pred, T(g['cot']))
    close(gp, g['g_pred'], 1e-9)
    val2 = L.lpips_vgg(pred, T(target_np), cw, cb, lins)
    kd = 3.0 * val2.mean()
    close(kd.detach(), g['kd_lpips'])
    g2, = torch.autograd.grad(kd, pred)
    close(g2, g['kd_lpips_g'], 1e-9)


@pytest.mark.parametrize('tag,size', [('s256', 256), ('s1024', 1024), ('s64', 64)])
def test_mask_glue_oracle_matches_reference(golden_dir, tag, size):
    """oracle.lpips_oracle.batch_img_preprocess / content_mask against Batch_Img_Parsing / Get_Masked_Tensor run
    unmodified (Util/content_aware_pruning.py:61-117)."""
    import sys
    sys.path.insert(0, golden_dir)
    import synth
    from oracle import lpips_oracle as L
    g = np.load(os.path.join(golden_dir, 'mask_glue.npz'))
    n = 2
    rs = np.random.RandomState(910 + size)
    img = torch.from_numpy((rs.standard_normal((n, 3, size, size)) * 0.8).astype(np.float32))
    pre = L.batch_img_preprocess(img)
    close(pre[:, :, ::7, ::5], g[f'{tag}.pre'], 1e-6)
    scores = torch.from_numpy(synth.parser_scores(911 + size, n))
    parsing = scores.argmax(1)
    assert int(parsing.sum()) == int(g[f'{tag}.parsing_sum'])
    mask = L.content_mask(parsing, size)
    want = np.unpackbits(g[f'{tag}.mask'])[:n * size * size].reshape(n, size, size)
    assert np.array_equal(mask.numpy().astype(np.uint8), want)
    assert 0.05 < want.mean() < 0.999
    masked = L.get_masked_tensor(img.double(), parsing)
    assert abs(float(masked.abs().sum()) - float(g[f'{tag}.masked_sum'])) <= 1e-6 * float(g[f'{tag}.masked_sum'])


@pytest.mark.parametrize('mode', ['Output_Only', 'Intermediate'])
def test_kd_step_full_loss_golden(golden_dir, mode):
    """oracle.kd_step with the parsed content mask and LPIPS against the fixture produced by the reference's own KD_loss
    text (train.py:145-184) + lpips.PerceptualLoss + Batch_Img_Parsing / Get_Masked_Tensor."""
    import synth
    g0 = np.load(os.path.join(golden_dir, 'kd_tiny.npz'))
    g = np.load(os.path.join(golden_dir, 'kd_full_tiny.npz'))
    c, dp, sp, tp, z, s_noise, t_noise, _ = _kd_tiny_inputs(g0)
    cw, cb = synth.vgg16_weights(synth.KD_FULL['seed_vgg'])
    lp = ([T(a) for a in cw], [T(a) for a in cb], [T(g[f'lin{k}']) for k in range(5)])
    parsing = torch.from_numpy(synth.parser_scores(synth.KD_FULL['seed_parser'], c['batch'])).argmax(1)
    loss, grads = O.kd_step(sp, tp, dp, c['size'], z, s_noise, t_noise, c['inject'], kd_mode=mode, parsing=parsing, lpips=lp)
    close(loss, g[f'{mode}.g_loss'] + g[f'{mode}.kd_l1'] + g[f'{mode}.kd_lpips'])
    assert float(g[f'{mode}.kd_lpips']) > 1e-3          # the LPIPS term is a real part of the loss in this fixture
    for n in g0['param_names']:
        ref = g[f'{mode}.grad.{n}']
        got = grads[n].numpy() if n in grads else np.zeros_like(ref)
        if np.abs(ref).max() == 0:
            assert np.abs(got).max() == 0, n
        else:
            close(got, ref, 1e-9)
