"""Seed-regenerated synthetic inputs shared by tests/golden/make_golden.py (which feeds them to the
UNMODIFIED reference) and by the tests (which feed them to the oracle / the CUDA path).

Large fixtures (a 30 M-parameter 256px generator, a 21 M-parameter discriminator, per-layer noise maps)
cannot be committed; they are re-drawn on both sides from numpy's frozen legacy `RandomState` stream
(bit-stable by NumPy policy), in the state_dict's key order.  Distributions follow the reference's
initialisers (model.py:399-521, :740-798) except that the zero-initialised small parameters
(noise.weight, activate.bias, ToRGB bias, mapping biases) are non-zero, because zeros hide bugs.
"""
import numpy as np

# fixture configurations (make_golden.py generates with them, the tests regenerate the inputs from them)
TINY = dict(size=32, style_dim=32, n_mlp=2, net_shape=[12, 12, 12, 12, 8, 8, 6, 6])
KD_TINY = dict(size=32, style_dim=32, n_mlp=2, student=[12, 12, 12, 12, 8, 8, 6, 6], teacher=[24] * 8, batch=4,
               inject=3, seed_student=31, seed_teacher=32, seed_disc=33, seed_inputs=34)
# BASELINE.json configs[2]: saliency of the full 256px generator over 64 latents, 8 batches of 8
CONFIG3 = dict(size=256, n_sample=64, batch_size=8, noise_prob=0.05, seed_weights=7, seed_batches=1000)


def _scale_for(key: str, shape) -> tuple:
    """(offset, gain) of the N(0,1) draw for a state_dict key."""
    if key.startswith('style.') and key.endswith('.weight'):
        return 0.0, 100.0                 # randn / lr_mul, lr_mul = 0.01 (model.py:144)
    if key.startswith('style.') and key.endswith('.bias'):
        return 0.0, 0.1
    if key.endswith('modulation.bias'):
        return 1.0, 0.1                   # bias_init = 1 (model.py:231)
    if key.endswith('noise.weight') or key.endswith('activate.bias'):
        return 0.0, 0.3
    if key.endswith('.bias') and len(shape) == 4:     # ToRGB bias [1,3,1,1]
        return 0.0, 0.3
    if key.endswith('.bias'):                          # discriminator biases
        return 0.0, 0.2
    return 0.0, 1.0


def synth_state(template, seed: int):
    """template: ordered {key: tensor-or-array with .shape} (a module's state_dict).  Returns
    {key: float64 ndarray} for every key except FIR buffers ('kernel'), which keep their values."""
    rs = np.random.RandomState(seed)
    out = {}
    for key, ref in template.items():
        shape = tuple(ref.shape)
        if key.endswith('kernel'):
            continue
        off, gain = _scale_for(key, shape)
        # fp32-representable values: the fp64 reference and the fp32 CUDA path see identical numbers
        out[key] = (off + gain * rs.standard_normal(shape)).astype(np.float32).astype(np.float64)
    return out


def load_synth(module, seed: int):
    """Overwrite a module's parameters / noise buffers with synth_state(seed) (dtype of the module)."""
    import torch
    sd = module.state_dict()
    new = synth_state(sd, seed)
    with torch.no_grad():
        for k, v in new.items():
            sd[k].copy_(torch.from_numpy(v).to(sd[k].dtype))
    return module


def latents_and_noise(seed: int, batch: int, latent_dim: int, noise_shapes, n_latents: int = 1):
    """fp32-representable latents and per-layer noise maps of one batch (float64 arrays holding fp32
    values, so that the fp64 reference and the fp32 CUDA path see identical numbers)."""
    rs = np.random.RandomState(seed)
    z = [rs.standard_normal((batch, latent_dim)).astype(np.float32).astype(np.float64) for _ in range(n_latents)]
    noise = [rs.standard_normal((batch, 1, h, w)).astype(np.float32).astype(np.float64) for (h, w) in noise_shapes]
    return z, noise, rs


def ellipse_mask(size: int) -> np.ndarray:
    """Synthetic stand-in for the BiSeNet face mask: centred ellipse."""
    yy, xx = np.mgrid[0:size, 0:size]
    c = (size - 1) / 2
    return (((yy - c) / (0.42 * size)) ** 2 + ((xx - c) / (0.34 * size)) ** 2) <= 1


# ------------------------------------------------------------------ LPIPS-VGG16 / content-mask fixtures
LPIPS_TINY = dict(size=32, batch=2, seed_weights=51, seed_inputs=52)
KD_FULL = dict(seed_vgg=53, seed_parser=54)     # on top of KD_TINY: the complete generator loss (LPIPS + parsed mask)
VGG16_CHANNELS = (64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512)


def vgg16_weights(seed: int):
    """The 13 convolutions of torchvision's vgg16().features[0:30], He-scaled N(0, 2/(9 I)) weights and N(0, 0.05^2) biases
    (the pretrained ones cannot be downloaded; activations keep a sane range through 13 layers)."""
    rs = np.random.RandomState(seed)
    ws, bs, cin = [], [], 3
    for cout in VGG16_CHANNELS:
        ws.append((rs.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (9 * cin))).astype(np.float32).astype(np.float64))
        bs.append((rs.standard_normal((cout,)) * 0.05).astype(np.float32).astype(np.float64))
        cin = cout
    return ws, bs


def lpips_images(seed: int, batch: int, size: int):
    """A 'student' and a 'teacher' image in [-1, 1]: smooth content plus a perturbation (pred = target + noise)."""
    rs = np.random.RandomState(seed)
    target = np.tanh(rs.standard_normal((batch, 3, size, size))).astype(np.float32).astype(np.float64)
    pred = np.clip(target + 0.3 * rs.standard_normal(target.shape), -1, 1).astype(np.float32).astype(np.float64)
    return pred, target


def parser_scores(seed: int, n: int, size: int = 512, classes: int = 19):
    """Stand-in face-parser output: smooth random class scores (blobs a few dozen pixels wide, so the mask has
    structure), float32 [n, classes, size, size]."""
    rs = np.random.RandomState(seed)
    coarse = rs.standard_normal((n, classes, size // 32, size // 32)).astype(np.float32)
    fine = np.repeat(np.repeat(coarse, 32, axis=2), 32, axis=3)
    return (fine + 0.35 * rs.standard_normal((n, classes, size, size)).astype(np.float32)).astype(np.float32)
