"""Per-layer CUDA-event timing of StyledConv forward on the KD-step shapes (development aid)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
import torch
import model
from b200gan import config

config.set_default_algo(config.ALGO_TCGEN05_TF32)
dev = 'cuda'
B = 16
layers = [(512, 512, 4, 0), (512, 512, 8, 0), (512, 512, 16, 0), (512, 512, 32, 0), (512, 512, 64, 0), (512, 256, 64, 1),
          (256, 256, 128, 0), (256, 128, 128, 1), (128, 128, 256, 0), (154, 154, 64, 0), (154, 77, 64, 1),
          (77, 77, 128, 0), (77, 39, 128, 1), (39, 39, 256, 0)]
prof = config.KernelProfiler()
for ci, co, r, up in layers:
    m = model.StyledConv(ci, co, 3, 512, upsample=bool(up)).to(dev)
    for p in m.parameters():
        p.requires_grad_(False)
    x = torch.randn(B, ci, r, r, device=dev)
    w = torch.randn(B, 512, device=dev)
    with torch.no_grad():
        for _ in range(2):
            m(x, w)
        torch.cuda.synchronize()
        prof.records.clear()
        config.set_profiler(prof)
        for _ in range(5):
            m(x, w)
        torch.cuda.synchronize()
        config.set_profiler(None)
    s = prof.summary()
    parts = []
    for k, v in s.items():
        tf = v['flops'] / (v['ms'] / 1e3) / 1e12 if v['flops'] else 0
        gb = v['bytes'] / (v['ms'] / 1e3) / 1e9
        parts.append(f"{k}: {v['ms'] / 5 * 1e3:7.1f} us {tf:6.1f} TF/s {gb:6.0f} GB/s")
    print(f'{ci:4d}->{co:4d} @{r:3d} up={up}: ' + ' | '.join(parts), flush=True)
