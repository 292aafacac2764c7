"""Development aid: compare the tcgen05 convolution against the fp32 SIMT engine, case by case."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
import torch
from b200gan import config, modconv as mc
from b200gan._lib import lib, check

torch.manual_seed(0)
dev = 'cuda'


def run(b, cin, cout, h, k, up=False, act=True):
    x = torch.randn(b, cin, h, h, device=dev)
    s = torch.nn.functional.pad(torch.randn(b, cin, device=dev) * 0.5 + 1, (0, mc.pitch_of(cin) - cin))
    w = torch.randn(1, cout, cin, k, k, device=dev)
    ho = h * 2 if up else h
    noise = torch.randn(b, 1, ho, ho, device=dev)
    nw = torch.tensor([0.3], device=dev)
    bias = torch.randn(cout, device=dev) * 0.2
    fir = (torch.tensor([1., 3., 3., 1.])[:, None] * torch.tensor([1., 3., 3., 1.])[None, :] / 64 * 4).to(dev)
    outs = []
    for algo in (0, 1):
        with config.use_algo(algo):
            y = mc.styled_conv(x, s, w, noise, nw, bias, 1.0 / (cin * k * k) ** 0.5, upsample=up, fir=fir, pad=(1, 1),
                               act=act)
        torch.cuda.synchronize()
        outs.append(y.float().clone())
    a, c = outs
    err = (a - c).abs().max() / a.abs().max()
    print(f'b={b} cin={cin} cout={cout} h={h} k={k} up={up}: rel err {err:.3e}', flush=True)
    return float(err)


cases = [(2, 32, 16, 8, 1), (2, 32, 32, 16, 1), (2, 64, 64, 16, 1), (2, 32, 32, 16, 3), (1, 40, 40, 16, 3),
         (3, 128, 128, 32, 3), (2, 160, 160, 16, 3), (16, 512, 512, 4, 3), (2, 512, 512, 8, 3), (2, 256, 256, 32, 3),
         (2, 80, 40, 16, 3, True), (2, 128, 128, 8, 3, True), (1, 64, 32, 5, 3)]
bad = 0
for c in cases:
    e = run(*c)
    bad += e > 3e-3
print('FAILED' if bad else 'ALL OK', bad)
