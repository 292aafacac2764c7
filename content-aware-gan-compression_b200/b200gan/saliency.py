"""Content-aware channel saliency (reference Util/content_aware_pruning.py:152-249, prune.py:39-56) on
the fp32-exact engines, with the multi-GPU sharding of SURVEY.md §8e.

    score_l[i] = mean_{o,ky,kx} | d/dW_l[0,o,i,ky,kx]  sum |noisy(img) - img| |      per batch,
    summed over batches afterwards (the |.| is taken AFTER the in-batch sum, so whole batches stay on a rank).

The face parser (BiSeNet) that produces the content mask is a third-party network outside the hot
path: the caller passes `mask_fn(image[1,3,S,S]) -> bool ndarray [S,S]` (default: a centred ellipse).
Every convolution of the pass runs in `config.exact_fp32()` (SIMT fp32, fixed reduction order):
the product of the pass -- the prune mask -- has to be reproducible bit for bit.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import config
from . import dist as D


def default_mask(img: torch.Tensor) -> np.ndarray:
    s = img.shape[-1]
    yy, xx = np.mgrid[0:s, 0:s]
    c = (s - 1) / 2
    return (((yy - c) / (0.42 * s)) ** 2 + ((xx - c) / (0.34 * s)) ** 2) <= 1


def salt_pepper_maps(mask: np.ndarray, prob: float, rng=np.random):
    """Randomness of Get_Salt_Pepper_Noisy_Image (reference :152-171), drawn in the same order: one
    randint(0, 2, (S, S)) field, then one uniform per *masked* pixel in raster order (the reference's
    `mask[h, w] == True and np.random.random() < prob` short-circuits on unmasked pixels)."""
    s = mask.shape[0]
    value = rng.randint(low=0, high=2, size=(s, s)) * 2 - 1
    hit = np.zeros((s, s), dtype=bool)
    n = int(mask.sum())
    hit[mask.astype(bool)] = rng.random_sample(n) < prob if hasattr(rng, 'random_sample') else rng.random(n) < prob
    return value.astype(np.float32), hit


def noisy_images(img: torch.Tensor, mask_fn: Callable, prob: float, rng=np.random) -> torch.Tensor:
    """Clone of `img` with +-1 written (to all channels) at the selected pixels of each sample."""
    vals, hits = [], []
    det = img.detach()
    for i in range(img.shape[0]):
        mask = np.asarray(mask_fn(det[i:i + 1]), dtype=bool)
        v, h = salt_pepper_maps(mask, prob, rng)
        vals.append(v)
        hits.append(h)
    val = torch.from_numpy(np.stack(vals)).to(img.device).unsqueeze(1)
    hit = torch.from_numpy(np.stack(hits)).to(img.device).unsqueeze(1)
    return torch.where(hit.expand_as(det), val.expand_as(det), det)


def _unwrap(generator):
    return generator.module if hasattr(generator, 'module') and not hasattr(generator, 'conv1') else generator


def weight_gradient_scores(generator) -> List[np.ndarray]:
    """Get_Weight_Gradient's read-out (reference :187-196): [conv1] + convs + [to_rgbs[-1]]."""
    g = _unwrap(generator)
    mods = [g.conv1] + list(g.convs) + [g.to_rgbs[-1]]
    return [m.conv.weight.grad.abs().mean(dim=(0, 1, 3, 4)).cpu().numpy() for m in mods]


def batch_sizes(n_sample: int, batch_size: int) -> List[int]:
    """Reference :217-219: the remainder is folded into the last batch."""
    n_batch = n_sample // batch_size
    return [batch_size] * (n_batch - 1) + [batch_size + n_sample % batch_size]


def content_aware_scores(generator, n_sample: int, batch_size: int, noise_prob: float, device,
                         mask_fn: Optional[Callable] = None, seed: Optional[int] = None,
                         latent_dim: int = 512) -> List[List[np.ndarray]]:
    """Get_Content_Aware_Pruning_Score (reference :200-249).  Returns, on every rank, the list over
    batches (in batch order) of per-layer score vectors.

    seed=None: consume the global torch / numpy RNG streams exactly like the reference (single process).
    seed=int : every batch owns its RNG stream (numpy RandomState seeded with seed+batch: latents, per-layer
               noise, salt & pepper, in that order), which makes the result independent of how batches are
               sharded over ranks and reproducible on any host (tests/golden/config3_saliency.npz).
    """
    mask_fn = mask_fn or default_mask
    sizes = batch_sizes(n_sample, batch_size)
    world = D.get_world_size()
    if world > 1 and seed is None:
        raise ValueError('multi-rank saliency needs an explicit seed (per-batch RNG streams)')
    mine = D.shard_batches(len(sizes)) if world > 1 else list(range(len(sizes)))
    g = _unwrap(generator)
    local: Dict[int, List[np.ndarray]] = {}
    with config.exact_fp32():
        for idx in mine:
            b = sizes[idx]
            if seed is None:
                z = torch.randn(b, latent_dim).to(device)
                noise, rng = None, np.random
            else:
                # one numpy RandomState per batch (frozen, platform-independent stream): latents, then the per-layer
                # noise maps in execution order, then the salt & pepper draws continue the same stream
                rng = np.random.RandomState(seed + idx)
                z = torch.from_numpy(rng.standard_normal((b, latent_dim)).astype(np.float32)).to(device)
                noise = [torch.from_numpy(rng.standard_normal((b, 1, n.shape[2], n.shape[3])).astype(np.float32)).to(device)
                         for n in g.make_noise()]
            generator.zero_grad()
            img = generator([z], noise=noise) if noise is not None else generator([z])
            noisy = noisy_images(img, mask_fn, noise_prob, rng)
            torch.sum(torch.abs(noisy - img)).backward()
            local[idx] = weight_gradient_scores(generator)
            generator.zero_grad()
    if world > 1:
        return D.gather_scores_in_batch_order(local, len(sizes))
    return [local[i] for i in range(len(sizes))]


def total_scores(per_batch: Sequence[Sequence[np.ndarray]]) -> List[np.ndarray]:
    """prune.py:45-46: sum over batches, layer by layer, in batch order."""
    n_layers = len(per_batch[0])
    out = []
    for layer in range(n_layers):
        acc = np.zeros_like(per_batch[0][layer])
        for b in per_batch:
            acc = acc + b[layer]
        out.append(acc)
    return out


def prune_masks(scores: Sequence[np.ndarray], remove_ratio: float) -> List[np.ndarray]:
    """Keep-masks: drop the int(C*ratio) lowest-score channels of each layer
    (Util/pruning_util.py:197-244, uniform removal list)."""
    masks = []
    for sc in scores:
        c = len(sc)
        keep = np.ones(c, dtype=bool)
        rm = int(c * remove_ratio)
        if 0 < rm < c:
            keep[np.argsort(sc)[:rm]] = False
        masks.append(keep)
    return masks
