"""BASELINE.json configs 3 and 5 timed on one GPU (reported in profiles/, not bench lines):
   config 3: content-aware saliency, full 256px generator, 64 latents in 8 batches of 8, exact-fp32 engines;
   config 5: get_fid.py-style sampling loop (Evaluation/fid.py:19-38), 256px pruned generator, batch 64."""
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import model
from b200gan import config, saliency

dev = torch.device('cuda')
out = {}
torch.manual_seed(0)
np.random.seed(0)
g = model.Generator(256, 512, 8).to(dev)
for algo_name, ctx in (('exact_fp32 (SIMT, the prune-mask path)', config.exact_fp32),
                       ('tcgen05_tf32', lambda: config.use_algo(config.ALGO_TCGEN05_TF32))):
    # content_aware_scores always enters exact_fp32(); time the tensor-pipe variant by patching the context
    orig = config.exact_fp32
    config.exact_fp32 = ctx
    saliency.config.exact_fp32 = ctx
    try:
        saliency.content_aware_scores(g, 16, 8, 0.05, dev, seed=1)        # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        per_batch = saliency.content_aware_scores(g, 64, 8, 0.05, dev, seed=1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    finally:
        config.exact_fp32 = orig
        saliency.config.exact_fp32 = orig
    masks = saliency.prune_masks(saliency.total_scores(per_batch), 0.7)
    out[f'saliency_64_latents[{algo_name}]'] = {'seconds': dt, 'images_per_s': 64 / dt,
                                                'tflops_ref_convention': 64 * 3 * 90.236e9 / dt / 1e12,
                                                'kept_channels': [int(m.sum()) for m in masks]}
    if 'exact' in algo_name:
        ref_masks = masks
    else:
        out['prune_mask_agreement_tf32_vs_fp32'] = float(np.mean([np.mean(a == b) for a, b in zip(masks, ref_masks)]))

config.set_default_algo(config.best_available_algo())
student = model.Generator(256, 512, 8, generator_net_shape=bench.STUDENT_SHAPES[256]).to(dev).eval()
with torch.no_grad():
    for _ in range(3):
        student([torch.randn(64, 512, device=dev)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 40
    e0.record()
    for _ in range(n):
        student([torch.randn(64, 512, device=dev)], truncation=1)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
out['fid_sampling_pruned_256_batch64'] = {'ms_per_batch': ms, 'images_per_s': 64 / (ms / 1e3),
                                          'note': 'generator only (Inception is a third-party network), eager launches'}
print(json.dumps(out))
