"""Isolated launches of the hot kernels on KD-step layer shapes (for `ncu --set full`, 1 GPU)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
import torch
import model
from b200gan import config

config.set_default_algo(config.ALGO_TCGEN05_TF32)
dev = 'cuda'
torch.manual_seed(0)
B = 16
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
# (tag, cin, cout, res_in, upsample, backward)
layers = [('teacher_512_64', 512, 512, 64, False, False), ('teacher_128_256', 128, 128, 256, False, False),
          ('teacher_up_256_128', 256, 128, 128, True, False), ('student_154_64', 154, 154, 64, False, True),
          ('student_39_256', 39, 39, 256, False, True), ('student_up_77_39', 77, 39, 128, True, True)]
mods = []
for tag, ci, co, r, up, bwd in layers:
    if which != 'all' and which not in tag:
        continue
    m = model.StyledConv(ci, co, 3, 512, upsample=up).to(dev)
    x = torch.randn(B, ci, r, r, device=dev, requires_grad=bwd)
    w = torch.randn(B, 512, device=dev)
    mods.append((tag, m, x, w, bwd))
for tag, m, x, w, bwd in mods:           # warm-up (attribute setup, caches)
    y = m(x, w)
    if bwd:
        y.sum().backward()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for tag, m, x, w, bwd in mods:
    y = m(x, w)
    if bwd:
        y.square().mean().backward()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
