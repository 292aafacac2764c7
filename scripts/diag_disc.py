"""Diagnostic: where does the input gradient of the 32px discriminator differ from fp64?  (development aid)"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'content-aware-gan-compression_b200'), ROOT, os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
os.environ.setdefault('CAGC_CONV_ALGO', 'simt')
import model, synth
from b200gan import config
from oracle import stylegan2_oracle as O
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
c = synth.KD_TINY
g = np.load(os.path.join(ROOT, 'tests/golden/kd_tiny.npz'))
disc = synth.load_synth(model.Discriminator(32), c['seed_disc']).cuda()
sd64 = {k: v.double() for k, v in disc.state_dict().items()}
x64 = torch.from_numpy(g['d_x']).cuda()
cot64 = torch.from_numpy(g['d_cot']).cuda()
ref_gx = torch.from_numpy(g['d_gx']).cuda()
def rel(a, b): return float((a.double() - b).abs().max() / b.abs().max())
# (1) oracle in fp64 ON THE GPU, from the fp32-rounded input
for name, xin in (('fp64 input', x64), ('fp32-rounded input', x64.float().double())):
    xr = xin.clone().requires_grad_(True)
    pred = O.discriminator_forward(sd64, xr, 32)
    gx, = torch.autograd.grad(pred, xr, cot64)
    print(f'oracle fp64 on GPU, {name}: pred {rel(pred, torch.from_numpy(g["d_pred"]).cuda()):.2e}  gx {rel(gx, ref_gx):.2e}')
# (2) oracle in fp32 on the GPU
sd32 = {k: v.float() for k, v in sd64.items()}
xr = x64.float().requires_grad_(True)
pred = O.discriminator_forward(sd32, xr, 32)
gx, = torch.autograd.grad(pred, xr, cot64.float())
print(f'oracle fp32 on GPU (library ops): gx {rel(gx, ref_gx):.2e}')
# (3) our module, frozen (own engines) and trainable (library convs + our ops)
for frozen in (True, False):
    for p in disc.parameters():
        p.requires_grad_(not frozen)
    xr = x64.float().requires_grad_(True)
    with config.exact_fp32():
        pred = disc(xr)
        gx, = torch.autograd.grad(pred, xr, cot64.float())
    e = (gx.double() - ref_gx).abs() / ref_gx.abs().max()
    idx = np.unravel_index(int(e.argmax()), e.shape)
    print(f'module frozen={frozen}: pred {rel(pred, torch.from_numpy(g["d_pred"]).cuda()):.2e} gx max {float(e.max()):.2e} at {idx}; '
          f'elements > 1e-4: {int((e > 1e-4).sum())} of {e.numel()}; per-sample max {[float(e[i].max()) for i in range(e.shape[0])]}')
