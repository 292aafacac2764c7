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
TINY = dict(size=32, style_dim=32, n_mlp=2, net_shape=[12, 12, 12, 12, 8, 8, 6, 6])


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


if __name__ == '__main__':
    gen_upfirdn2d()
    gen_fused_act()
    gen_layers()
    gen_generator()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)))
