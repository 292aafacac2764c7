#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference on CPU in fp64.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference's `op/*.py` JIT-compile a CUDA extension at import
(op/fused_act.py:11-17, op/upfirdn2d.py:10-16); on CPU the extension is never
called (native fallbacks op/fused_act.py:105-116, op/upfirdn2d.py:146-149), so
`torch.utils.cpp_extension.load` is stubbed to skip the 100 s build.  Nothing
else of the reference is altered.  The fixtures written here (small .npz
files) are what travels to the GPU box; /root/reference does not.
"""
import os
import random
import sys

import numpy as np
import torch

REF = os.environ.get('CAGC_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))

import torch.utils.cpp_extension as _ce  # noqa: E402

_ce.load = lambda *a, **k: None
sys.path.insert(0, REF)
import model as ref_model  # noqa: E402
from op import upfirdn2d as ref_upfirdn2d, fused_leaky_relu as ref_fused_leaky_relu  # noqa: E402
from Util.network_util import Get_Network_Shape, Build_Generator_From_Dict  # noqa: E402
from Util.mask_util import Mask_the_Generator  # noqa: E402
from Util.pruning_util import Get_Uniform_RmveList, Generate_Prune_Mask_List  # noqa: E402

torch.set_default_dtype(torch.float64)


def seed(s):
    torch.manual_seed(s)
    np.random.seed(s)
    random.seed(s)


def npy(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------ upfirdn2d
def gen_upfirdn2d():
    seed(1)
    out = {}
    cases = [
        # (n, c, h, w, taps or 2-D kernel size, up, down, pad0, pad1, gain)
        (2, 3, 9, 9, [1, 3, 3, 1], 1, 1, 1, 1, 4.0),      # Blur after up-conv (model.py:207-213)
        (2, 3, 8, 8, [1, 3, 3, 1], 2, 1, 2, 1, 4.0),      # Upsample in ToRGB (model.py:38-56)
        (1, 2, 16, 16, [1, 3, 3, 1], 1, 2, 1, 1, 1.0),    # Downsample (model.py:59-77)
        (2, 2, 8, 8, [1, 3, 3, 1], 1, 1, 2, 2, 1.0),      # Blur before stride-2 conv in D (model.py:683-689)
        (1, 3, 7, 5, [1, 2, 1], 1, 1, 1, 1, 1.0),         # odd sizes, 3-tap
        (1, 2, 6, 6, [1, 3, 3, 1], 1, 1, -1, 2, 1.0),     # negative pad (crop)
        (1, 2, 6, 7, [1, 3, 3, 1], 2, 2, 1, -1, 1.0),     # up and down together, crop at the end
        (1, 1, 5, 5, 'rand5x3', 1, 1, 2, 1, 1.0),         # non-separable, non-square kernel
        (1, 2, 4, 4, [1, 3, 3, 1], 3, 2, 2, 2, 9.0),      # up 3 / down 2
        (2, 2, 33, 33, [1, 3, 3, 1], 1, 1, 1, 1, 4.0),    # 33^2 -> 32^2 as in the generator
    ]
    out['n_cases'] = np.array(len(cases))
    for i, (n, c, h, w, taps, up, down, p0, p1, gain) in enumerate(cases):
        if taps == 'rand5x3':
            k = torch.randn(5, 3)
        else:
            k = ref_model.make_kernel(taps).double() * gain
        x = torch.randn(n, c, h, w, requires_grad=True)
        y = ref_upfirdn2d(x, k, up=up, down=down, pad=(p0, p1))
        gy = torch.randn_like(y)
        gx, = torch.autograd.grad(y, x, gy)
        out[f'c{i}_x'] = npy(x)
        out[f'c{i}_k'] = npy(k)
        out[f'c{i}_cfg'] = np.array([up, down, p0, p1])
        out[f'c{i}_y'] = npy(y)
        out[f'c{i}_gy'] = npy(gy)
        out[f'c{i}_gx'] = npy(gx)
    np.savez_compressed(os.path.join(HERE, 'upfirdn2d.npz'), **out)


# ------------------------------------------------------------------ fused act
def gen_fused_act():
    seed(2)
    out = {}
    shapes = [(4, 6), (2, 5, 7, 3), (3, 8, 4, 4)]
    out['n_cases'] = np.array(len(shapes))
    for i, shp in enumerate(shapes):
        x = torch.randn(*shp, requires_grad=True)
        b = torch.randn(shp[1], requires_grad=True)
        y = ref_fused_leaky_relu(x, b)
        gy = torch.randn_like(y)
        gx, gb = torch.autograd.grad(y, [x, b], gy)
        out[f'c{i}_x'] = npy(x)
        out[f'c{i}_b'] = npy(b)
        out[f'c{i}_y'] = npy(y)
        out[f'c{i}_gy'] = npy(gy)
        out[f'c{i}_gx'] = npy(gx)
        out[f'c{i}_gb'] = npy(gb)
    x = torch.randn(3, 4, 5)
    out['nobias_x'] = npy(x)
    out['nobias_y'] = npy(ref_fused_leaky_relu(x))
    np.savez_compressed(os.path.join(HERE, 'fused_act.npz'), **out)


# ------------------------------------------------------------------ layers
def randomize_small_params(mod):
    """noise.weight / activate.bias / ToRGB bias are zero-initialised in the reference
    (model.py:296,378; op/fused_act.py:92) which hides bugs: redraw them."""
    for name, p in mod.named_parameters():
        if name.endswith('noise.weight') or name.endswith('activate.bias') or name.endswith('to_rgb1.bias') \
                or (name.endswith('.bias') and p.ndim == 4) or name == 'bias':
            with torch.no_grad():
                p.copy_(torch.randn_like(p) * 0.5)


def gen_layers():
    seed(3)
    out = {}
    style_dim = 16
    cfgs = [('same', 10, 7, False, 6), ('up', 10, 7, True, 5), ('same39', 39, 20, False, 8)]
    for tag, cin, cout, up, res in cfgs:
        m = ref_model.StyledConv(cin, cout, 3, style_dim, upsample=up).double()
        randomize_small_params(m)
        x = torch.randn(3, cin, res, res, requires_grad=True)
        w = torch.randn(3, style_dim, requires_grad=True)
        ores = res * 2 if up else res
        nz = torch.randn(3, 1, ores, ores)
        y = m(x, w, noise=nz)
        gy = torch.randn_like(y)
        params = dict(m.named_parameters())
        names = sorted(params)
        grads = torch.autograd.grad(y, [x, w] + [params[n] for n in names], gy)
        for k, v in m.state_dict().items():
            out[f'{tag}.sd.{k}'] = npy(v)
        out[f'{tag}.x'] = npy(x)
        out[f'{tag}.w'] = npy(w)
        out[f'{tag}.noise'] = npy(nz)
        out[f'{tag}.y'] = npy(y)
        out[f'{tag}.gy'] = npy(gy)
        out[f'{tag}.gx'] = npy(grads[0])
        out[f'{tag}.gw'] = npy(grads[1])
        for n, g in zip(names, grads[2:]):
            out[f'{tag}.grad.{n}'] = npy(g)
    # ToRGB with skip
    m = ref_model.ToRGB(10, style_dim).double()
    randomize_small_params(m)
    x = torch.randn(2, 10, 8, 8, requires_grad=True)
    w = torch.randn(2, style_dim, requires_grad=True)
    skip = torch.randn(2, 3, 4, 4, requires_grad=True)
    y = m(x, w, skip)
    gy = torch.randn_like(y)
    params = dict(m.named_parameters())
    names = sorted(params)
    grads = torch.autograd.grad(y, [x, w, skip] + [params[n] for n in names], gy)
    for k, v in m.state_dict().items():
        out[f'rgb.sd.{k}'] = npy(v)
    out['rgb.x'], out['rgb.w'], out['rgb.skip'] = npy(x), npy(w), npy(skip)
    out['rgb.y'], out['rgb.gy'] = npy(y), npy(gy)
    out['rgb.gx'], out['rgb.gw'], out['rgb.gskip'] = npy(grads[0]), npy(grads[1]), npy(grads[2])
    for n, g in zip(names, grads[3:]):
        out[f'rgb.grad.{n}'] = npy(g)
    np.savez_compressed(os.path.join(HERE, 'layers.npz'), **out)


# ------------------------------------------------------------------ tiny generator
sys.path.insert(0, HERE)
from synth import TINY, KD_TINY, CONFIG3  # noqa: E402  (shared with the tests)


def salt_pepper_fn(mask, noise_pm1, hit):
    """Vectorised equivalent of Get_Salt_Pepper_Noisy_Image (Util/content_aware_pruning.py:152-171)
    for pre-drawn randomness: where mask & hit, every channel of the pixel is set to +-1."""
    sel = torch.from_numpy(mask & hit)
    val = torch.from_numpy(noise_pm1.astype(np.float64))

    def fn(img):
        out = img.clone()
        sel_b = sel.view(1, 1, *sel.shape).expand_as(out)
        return torch.where(sel_b, val.view(1, 1, *val.shape).expand_as(out), out)
    return fn


def gen_generator():
    seed(4)
    out = {}
    g = ref_model.Generator(TINY['size'], TINY['style_dim'], TINY['n_mlp'],
                            generator_net_shape=TINY['net_shape']).double()
    randomize_small_params(g)
    sd = g.state_dict()
    for k, v in sd.items():
        out[f'sd.{k}'] = npy(v)
    b = 3
    z1, z2 = torch.randn(b, TINY['style_dim']), torch.randn(b, TINY['style_dim'])
    noise = [torch.randn(b, 1, n.shape[2], n.shape[3]) for n in g.make_noise()]
    out['z1'], out['z2'] = npy(z1), npy(z2)
    for i, n in enumerate(noise):
        out[f'noise{i}'] = npy(n)
    # (a) single latent, explicit noise
    img = g([z1], noise=noise)
    out['img_single'] = npy(img)
    # (b) style mixing, rgb list, grads of every parameter for a fixed cotangent on the last image
    inject = 3
    out['inject_index'] = np.array(inject)
    rgbs = g([z1, z2], inject_index=inject, noise=noise, return_rgb_list=True)
    for i, r in enumerate(rgbs):
        out[f'rgb{i}'] = npy(r)
    cot = torch.randn_like(rgbs[-1])
    out['cot'] = npy(cot)
    params = dict(g.named_parameters())
    names = sorted(params)
    loss = (rgbs[-1] * cot).sum() + sum((r * r).mean() for r in rgbs[:-1])
    grads = torch.autograd.grad(loss, [params[n] for n in names])
    for n, gr in zip(names, grads):
        out[f'grad.{n}'] = npy(gr)
    # (c) stored noise buffers (randomize_noise=False) and truncation / input_is_latent
    img_fixed = g([z1], randomize_noise=False)
    out['img_fixed_noise'] = npy(img_fixed)
    mean_w = g.style(torch.randn(64, TINY['style_dim'])).mean(0, keepdim=True)
    out['mean_w'] = npy(mean_w)
    out['img_trunc'] = npy(g([z1], truncation=0.7, truncation_latent=mean_w, noise=noise))
    wlat = g.get_latent(z1)
    out['w_latent'] = npy(wlat)
    out['img_from_w'] = npy(g(None, input_is_latent=True, latent_styles=[wlat], noise=noise))
    # (d) saliency metric of one batch, exactly Get_Weight_Gradient (Util/content_aware_pruning.py:174-196)
    S = TINY['size']
    yy, xx = np.mgrid[0:S, 0:S]
    mask = (np.abs(yy - S / 2 + 0.5) < S * 0.3) & (np.abs(xx - S / 2 + 0.5) < S * 0.35)
    pm1 = np.random.randint(0, 2, size=(S, S)) * 2 - 1
    hit = np.random.random((S, S)) < 0.25
    out['sal_mask'], out['sal_pm1'], out['sal_hit'] = mask, pm1, hit
    g.zero_grad()
    img = g([z1], noise=noise)
    noisy = salt_pepper_fn(mask, pm1, hit)(img.detach())
    torch.sum(torch.abs(noisy - img)).backward()
    mods = [g.conv1] + list(g.convs) + [g.to_rgbs[-1]]
    scores = [m.conv.weight.grad.abs().mean(dim=(0, 1, 3, 4)).numpy() for m in mods]
    for i, s in enumerate(scores):
        out[f'score{i}'] = s
    # (e) prune with the reference's own bookkeeping and run the pruned net
    net_shape = Get_Network_Shape(sd)
    out['net_shape'] = np.array(net_shape)
    rmve = Get_Uniform_RmveList(net_shape, 0.5)
    masks = Generate_Prune_Mask_List(scores, net_shape, rmve)
    for i, m in enumerate(masks):
        out[f'prune_mask{i}'] = np.asarray(m)
    pruned = Mask_the_Generator(sd, masks)
    gp = Build_Generator_From_Dict(pruned, size=TINY['size'], latent=TINY['style_dim'], n_mlp=TINY['n_mlp']).double()
    out['pruned_net_shape'] = np.array(Get_Network_Shape(gp.state_dict()))
    for k, v in gp.state_dict().items():
        out[f'pruned_sd.{k}'] = npy(v)
    out['img_pruned'] = npy(gp([z1], noise=noise))
    out['sd_keys'] = np.array(list(sd.keys()))
    # (f) path-length regulariser quantity (model.py:661-666) with the noise drawn under a fixed seed
    torch.manual_seed(77)
    _, pl = g([z1], PPL_regularize=True, noise=noise)
    torch.manual_seed(77)
    pl_noise = torch.randn(b, 3, S, S) / np.sqrt(S * S)
    out['pl_noise'] = npy(pl_noise)
    out['path_lengths'] = npy(pl)
    np.savez_compressed(os.path.join(HERE, 'generator_tiny.npz'), **out)


# ------------------------------------------------------------------ discriminator + KD-like step
def gen_kd_tiny():
    """Reference `Discriminator` (model.py:740-798) forward + input gradient, and the KD-like generator step
    composed from the reference's own modules the way train.py:280-308 / :145-169 composes them (student
    forward with rgb list -> D -> g_nonsaturating_loss :215-218; teacher forward; content mask multiply
    (Get_Masked_Tensor's effect on a constant mask); L1 KD in both kd_modes).  Weights / inputs are re-drawn
    from seeds (tests/golden/synth.py): only the reference's OUTPUTS are stored."""
    import synth
    c = KD_TINY
    out = {}
    disc = synth.load_synth(ref_model.Discriminator(c['size']).double(), c['seed_disc'])
    student = synth.load_synth(ref_model.Generator(c['size'], c['style_dim'], c['n_mlp'],
                                                   generator_net_shape=c['student']).double(), c['seed_student'])
    teacher = synth.load_synth(ref_model.Generator(c['size'], c['style_dim'], c['n_mlp'],
                                                   generator_net_shape=c['teacher']).double(), c['seed_teacher'])
    shapes = [(n.shape[2], n.shape[3]) for n in student.make_noise()]
    z, noise2, rs = synth.latents_and_noise(c['seed_inputs'], c['batch'], c['style_dim'], shapes + shapes, n_latents=2)
    z = [torch.from_numpy(a) for a in z]
    s_noise = [torch.from_numpy(a) for a in noise2[:len(shapes)]]
    t_noise = [torch.from_numpy(a) for a in noise2[len(shapes):]]
    # (a) discriminator alone: logits and the gradient w.r.t. its input image
    x = torch.from_numpy(rs.standard_normal((c['batch'], 3, c['size'], c['size']))).requires_grad_(True)
    cot = torch.from_numpy(rs.standard_normal((c['batch'], 1)))
    pred = disc(x)
    gx, = torch.autograd.grad(pred, x, cot)
    out['d_x'], out['d_cot'], out['d_pred'], out['d_gx'] = npy(x), npy(cot), npy(pred), npy(gx)
    pred2 = disc(x[:2])                      # batch smaller than stddev_group (model.py:785)
    out['d_pred_b2'] = npy(pred2)
    # (b) KD-like step, both kd_modes (train.py:163-169)
    mask = torch.from_numpy(synth.ellipse_mask(c['size']).astype(np.float64)).view(1, 1, c['size'], c['size'])
    for p in disc.parameters():
        p.requires_grad_(False)              # train.py:287
    names = [n for n, _ in student.named_parameters()]
    for mode in ('Output_Only', 'Intermediate'):
        student.zero_grad()
        fake_list = student(z, return_rgb_list=True, inject_index=c['inject'], noise=s_noise)   # train.py:291
        g_loss = F.softplus(-disc(fake_list[-1])).mean()                                       # train.py:293-294, 215-218
        real_list = teacher(z, return_rgb_list=True, inject_index=c['inject'], noise=t_noise)   # train.py:151
        real_list = [r.detach() for r in real_list]
        if mode == 'Output_Only':
            kd = 3.0 * torch.mean(torch.abs(real_list[-1] * mask - fake_list[-1] * mask))       # train.py:157-164
        else:
            # train.py:165-169: the masked pair for the last image, the plain pairs are the loop's other items.  The
            # reference loop re-binds fake_img to the list entries, i.e. it uses the UNMASKED student images and the
            # unmasked teacher images for every resolution (fake_img_teacher_list is not masked): restated as such.
            kd = 3.0 * sum(torch.mean(torch.abs(r - f)) for r, f in zip(real_list, fake_list))
        total = g_loss + kd
        total.backward()
        out[f'{mode}.g_loss'], out[f'{mode}.kd'] = npy(g_loss), npy(kd)
        for n, p in student.named_parameters():
            out[f'{mode}.grad.{n}'] = npy(p.grad) if p.grad is not None else np.zeros(tuple(p.shape))
        if mode == 'Output_Only':
            out['fake_last'], out['real_last'] = npy(fake_list[-1]), npy(real_list[-1])
    out['param_names'] = np.array(names)
    out['d_keys'] = np.array(list(disc.state_dict().keys()))
    out['student_keys'] = np.array(list(student.state_dict().keys()))
    np.savez_compressed(os.path.join(HERE, 'kd_tiny.npz'), **out)


# ------------------------------------------------------------------ RNG consumption order
def gen_rng_order():
    """`randomize_noise=True`: NoiseInjection draws `image.new_empty(B,1,H,W).normal_()` per layer in call order
    (model.py:299-301).  Store the reference image produced by its INTERNAL draws under a fixed seed together
    with the noise list obtained by replaying the recipe (one normal_() per layer, execution order, output
    resolution) under the same seed -- and assert here that feeding that list explicitly reproduces the image."""
    import synth
    out = {}
    g = synth.load_synth(ref_model.Generator(TINY['size'], TINY['style_dim'], TINY['n_mlp'],
                                             generator_net_shape=TINY['net_shape']).double(), 41)
    b = 3
    z = torch.from_numpy(np.random.RandomState(42).standard_normal((b, TINY['style_dim'])))
    torch.manual_seed(99)
    img_internal = g([z])                                   # fresh noise drawn inside NoiseInjection
    torch.manual_seed(99)
    replay = [torch.empty(b, 1, n.shape[2], n.shape[3]).normal_() for n in g.make_noise()]
    # make_noise() itself consumed the stream; redo the replay with shapes only
    shapes = [(n.shape[2], n.shape[3]) for n in replay]
    torch.manual_seed(99)
    replay = [torch.empty(b, 1, h, w).normal_() for (h, w) in shapes]
    img_replay = g([z], noise=replay)
    assert torch.equal(img_internal, img_replay), 'replayed draw order differs from the reference internal order'
    out['z'], out['img'] = npy(z), npy(img_internal)
    for i, n in enumerate(replay):
        out[f'noise{i}'] = npy(n)
    out['seed_weights'] = np.array(41)
    np.savez_compressed(os.path.join(HERE, 'rng_order.npz'), **out)


# ------------------------------------------------------------------ BASELINE config 3: saliency at full size
def gen_config3(max_batches=None):
    """BASELINE.json configs[2]: content-aware saliency of the FULL 256px generator over 64 latents (8 batches of
    8).  The loop is Get_Content_Aware_Pruning_Score (Util/content_aware_pruning.py:217-247) with
    the reference's own Get_Salt_Pepper_Noisy_Image (:152-171) and Get_Weight_Gradient (:174-196), unmodified;
    the BiSeNet mask is replaced by the synthetic ellipse (no face in a random-init generator's output) and the
    per-layer noise is passed explicitly so that both sides see the same numbers.  fp64 on CPU: ~15 min here."""
    import time
    import synth
    from Util.content_aware_pruning import Get_Salt_Pepper_Noisy_Image, Get_Weight_Gradient
    c = CONFIG3
    g = synth.load_synth(ref_model.Generator(c['size'], 512, 8).double(), c['seed_weights'])
    shapes = [(n.shape[2], n.shape[3]) for n in g.make_noise()]
    mask = synth.ellipse_mask(c['size'])
    n_batch = c['n_sample'] // c['batch_size']
    sizes = [c['batch_size']] * (n_batch - 1) + [c['batch_size'] + c['n_sample'] % c['batch_size']]
    out = {'n_batches': np.array(len(sizes)), 'batch_sizes': np.array(sizes)}
    csum = 0.0
    for k, v in g.state_dict().items():
        csum += float(v.abs().sum())
    out['weights_abs_sum'] = np.array(csum)
    for idx, b in enumerate(sizes[:max_batches]):
        t0 = time.time()
        z, noise, rs = synth.latents_and_noise(c['seed_batches'] + idx, b, 512, shapes)
        np.random.set_state(rs.get_state())          # the salt & pepper draws continue the batch's stream
        img = g(noise_z=[torch.from_numpy(z[0])], noise=[torch.from_numpy(n) for n in noise])
        noisy = torch.cat([Get_Salt_Pepper_Noisy_Image(img[i:i + 1], mask, c['noise_prob']) for i in range(b)])
        scores = Get_Weight_Gradient(noisy, img, g)
        g.zero_grad()
        for li, s in enumerate(scores):
            out[f'b{idx}.score{li}'] = s
        out[f'b{idx}.img_abs_sum'] = np.array(float(img.detach().abs().sum()))
        print(f'config3 batch {idx}: {time.time() - t0:.1f} s', flush=True)
        np.savez_compressed(os.path.join(HERE, 'config3_saliency.npz'), **out)


# ------------------------------------------------------------------ LPIPS-VGG16 + content-mask glue (rest of the KD loss)
def _import_ref_lpips():
    """`import lpips` of the reference needs skimage / IPython at import time (lpips/__init__.py:7,
    lpips/networks_basic.py:11-12; none of it is used by the net-lin VGG distance) and downloads the torchvision
    VGG16 weights (lpips/pretrained_networks.py:100): empty stub modules, and a VGG16 that keeps its random init."""
    import types
    for name in ('skimage', 'skimage.measure', 'skimage.color', 'skimage.transform', 'IPython'):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.__path__ = []
                sys.modules[name] = m
    sys.modules['skimage.measure'].__dict__.setdefault('compare_ssim', None)
    sys.modules['IPython'].__dict__.setdefault('embed', lambda *a, **k: None)
    import torchvision.models as tvm
    real = tvm.vgg16
    tvm.vgg16 = lambda pretrained=True, **k: real(weights=None)
    import lpips as ref_lpips
    return ref_lpips


def gen_lpips():
    """The reference's own `lpips.PerceptualLoss(model='net-lin', net='vgg')` (train.py:510) on CPU in fp64: VGG16
    weights re-drawn from a seed (tests/golden/synth.py; torchvision's are not available offline), the five `lin`
    layers from the vendored lpips/weights/v0.1/vgg.pth.  Stores the distances and the gradient towards `pred`
    (the student image of train.py:182)."""
    import synth
    ref_lpips = _import_ref_lpips()
    c = synth.LPIPS_TINY
    loss = ref_lpips.PerceptualLoss(model='net-lin', net='vgg', use_gpu=False)
    net = loss.model.net.double()
    vgg = net.net
    convs = [m for sl in (vgg.slice1, vgg.slice2, vgg.slice3, vgg.slice4, vgg.slice5) for m in sl
             if isinstance(m, torch.nn.Conv2d)]
    cw, cb = synth.vgg16_weights(c['seed_weights'])
    with torch.no_grad():
        for m, wt, bt in zip(convs, cw, cb):
            m.weight.copy_(torch.from_numpy(wt))
            m.bias.copy_(torch.from_numpy(bt))
    assert not net.training
    out = {}
    for k, lin in enumerate((net.lin0, net.lin1, net.lin2, net.lin3, net.lin4)):
        out[f'lin{k}'] = npy(lin.model[1].weight.reshape(-1))          # vendored (7 KB): stored
    pred_np, target_np = synth.lpips_images(c['seed_inputs'], c['batch'], c['size'])
    pred = torch.from_numpy(pred_np).requires_grad_(True)
    target = torch.from_numpy(target_np)
    val = loss(pred, target)                                            # lpips/__init__.py:27-41
    cot = torch.from_numpy(np.random.RandomState(c['seed_inputs'] + 1).standard_normal(tuple(val.shape)))
    g, = torch.autograd.grad(val, pred, cot)
    out['val'], out['cot'], out['g_pred'] = npy(val), npy(cot), npy(g)
    # train.py:179-182 form: kd_lpips_lambda * mean(percept_loss(fake, teacher))
    val2 = loss(pred, target)
    g2, = torch.autograd.grad(3.0 * torch.mean(val2), pred)
    out['kd_lpips'], out['kd_lpips_g'] = npy(3.0 * torch.mean(val2)), npy(g2)
    np.savez_compressed(os.path.join(HERE, 'lpips_tiny.npz'), **out)


def gen_mask_glue():
    """`Batch_Img_Parsing` and `Get_Masked_Tensor` (Util/content_aware_pruning.py:61-117) run unmodified on CPU, with a
    stand-in parser that returns seeded class scores (BiSeNet's call contract: `parsing_net(x)[0]` = [N,19,512,512]); the
    tensor the parser RECEIVES is recorded (it is the preprocessing under test)."""
    import synth
    import Util.content_aware_pruning as cap
    out = {}
    for tag, size in (('s256', 256), ('s1024', 1024), ('s64', 64)):
        n = 2
        rs = np.random.RandomState(910 + size)
        img = torch.from_numpy((rs.standard_normal((n, 3, size, size)) * 0.8).astype(np.float32)).float()
        scores = synth.parser_scores(911 + size, n)
        seen = {}

        class Parser:
            def __call__(self, x):
                seen['x'] = x.detach().clone()
                return (torch.from_numpy(scores),)
        parsing = cap.Batch_Img_Parsing(img.clone(), Parser(), 'cpu')
        masked = cap.Get_Masked_Tensor(img.clone(), parsing, 'cpu', mask_grad=True)
        out[f'{tag}.pre'] = npy(seen['x'][:, :, ::7, ::5]).astype(np.float32)       # subsampled: 6 MB per case otherwise
        out[f'{tag}.parsing_sum'] = np.array(int(parsing.sum()))
        mask = (masked != 0).any(1)            # img has no exact zeros: the mask is where the product survived
        out[f'{tag}.mask'] = np.packbits(npy(mask).astype(np.uint8))
        out[f'{tag}.masked_sum'] = np.array(float(masked.detach().double().abs().sum()))
    np.savez_compressed(os.path.join(HERE, 'mask_glue.npz'), **out)


def _ref_train_functions(names):
    """Function definitions of the reference's train.py, executed from its own source text (train.py parses the command
    line and builds datasets at import, so it cannot be imported): {name: function}, globals = what train.py imports."""
    import ast
    src = open(os.path.join(REF, 'train.py')).read()
    tree = ast.parse(src)
    import train_hyperparams
    from Util.content_aware_pruning import Batch_Img_Parsing, Get_Masked_Tensor
    ns = {'torch': torch, 'F': F, 'np': np, 'train_hyperparams': train_hyperparams, 'device': 'cpu',
          'Batch_Img_Parsing': Batch_Img_Parsing, 'Get_Masked_Tensor': Get_Masked_Tensor}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), os.path.join(REF, 'train.py'), 'exec'), ns)
    return {n: ns[n] for n in names}


def gen_kd_full():
    """The COMPLETE generator loss of train.py:280-308: non-saturating GAN term + the reference's own `KD_loss`
    (train.py:145-184, executed from its source text) with its own lpips.PerceptualLoss, Batch_Img_Parsing and
    Get_Masked_Tensor; the face parser is a stand-in returning seeded class scores (BiSeNet's call contract), the VGG16
    weights are re-drawn from a seed.  Both kd_modes; every student gradient."""
    import argparse
    import synth
    ref_lpips = _import_ref_lpips()
    fns = _ref_train_functions(['KD_loss', 'Downsample_Image_256', 'g_nonsaturating_loss'])
    c = synth.KD_TINY
    out = {}
    disc = synth.load_synth(ref_model.Discriminator(c['size']).double(), c['seed_disc'])
    student = synth.load_synth(ref_model.Generator(c['size'], c['style_dim'], c['n_mlp'],
                                                   generator_net_shape=c['student']).double(), c['seed_student'])
    teacher = synth.load_synth(ref_model.Generator(c['size'], c['style_dim'], c['n_mlp'],
                                                   generator_net_shape=c['teacher']).double(), c['seed_teacher'])
    for p in list(disc.parameters()) + list(teacher.parameters()):
        p.requires_grad_(False)              # train.py:287, :507
    shapes = [(n.shape[2], n.shape[3]) for n in student.make_noise()]
    z, noise2, rs = synth.latents_and_noise(c['seed_inputs'], c['batch'], c['style_dim'], shapes + shapes, n_latents=2)
    z = [torch.from_numpy(a) for a in z]
    s_noise = [torch.from_numpy(a) for a in noise2[:len(shapes)]]
    t_noise = [torch.from_numpy(a) for a in noise2[len(shapes):]]
    loss = ref_lpips.PerceptualLoss(model='net-lin', net='vgg', use_gpu=False)
    net = loss.model.net.double()
    vgg = net.net
    convs = [m for sl in (vgg.slice1, vgg.slice2, vgg.slice3, vgg.slice4, vgg.slice5) for m in sl
             if isinstance(m, torch.nn.Conv2d)]
    cw, cb = synth.vgg16_weights(synth.KD_FULL['seed_vgg'])
    with torch.no_grad():
        for m, wt, bt in zip(convs, cw, cb):
            m.weight.copy_(torch.from_numpy(wt))
            m.bias.copy_(torch.from_numpy(bt))
    scores = torch.from_numpy(synth.parser_scores(synth.KD_FULL['seed_parser'], c['batch']))

    def parsing_net(x):
        return (scores,)

    def teacher_g(latents, **kw):            # train.py:151 draws fresh noise; the fixture pins it
        return teacher(latents, noise=t_noise, **kw)
    for mode in ('Output_Only', 'Intermediate'):
        args = argparse.Namespace(kd_mode=mode, kd_l1_lambda=3.0, kd_lpips_lambda=3.0, size=c['size'])
        student.zero_grad()
        fake_list = student(z, return_rgb_list=True, inject_index=c['inject'], noise=s_noise)           # train.py:291
        g_loss = fns['g_nonsaturating_loss'](disc(fake_list[-1]))                                      # train.py:293-294
        l1, lp = fns['KD_loss'](args, teacher_g, z, c['inject'], fake_list[-1], fake_list, loss, parsing_net)
        (g_loss + l1 + lp).backward()                                                                  # train.py:304-307
        out[f'{mode}.g_loss'], out[f'{mode}.kd_l1'], out[f'{mode}.kd_lpips'] = npy(g_loss), npy(l1), npy(lp)
        for n, p in student.named_parameters():
            out[f'{mode}.grad.{n}'] = npy(p.grad) if p.grad is not None else np.zeros(tuple(p.shape))
    for k, lin in enumerate((net.lin0, net.lin1, net.lin2, net.lin3, net.lin4)):
        out[f'lin{k}'] = npy(lin.model[1].weight.reshape(-1))
    np.savez_compressed(os.path.join(HERE, 'kd_full_tiny.npz'), **out)


if __name__ == '__main__':
    which = sys.argv[1:] or ['upfirdn2d', 'fused_act', 'layers', 'generator']
    sys.path.insert(0, HERE)
    from torch.nn import functional as F  # noqa: E402
    table = {'upfirdn2d': gen_upfirdn2d, 'fused_act': gen_fused_act, 'layers': gen_layers, 'generator': gen_generator,
             'kd_tiny': gen_kd_tiny, 'rng_order': gen_rng_order, 'config3': gen_config3, 'lpips': gen_lpips,
             'mask_glue': gen_mask_glue, 'kd_full': gen_kd_full}
    for w in which:
        table[w]()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)))
