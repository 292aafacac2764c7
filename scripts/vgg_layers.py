"""CUDA-event timing of the LPIPS-VGG16 convolutions (forward at 2N images, data gradient at N) -- development aid."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
import torch
from b200gan import config
from b200gan.dconv import _prep, _empty
from b200gan.lpips import _vconv
from b200gan._lib import stream_of

N = int(os.environ.get('BATCH', 16))
algo = config.ALGO_TCGEN05_TF32
layers = [(64, 64, 256), (64, 128, 128), (128, 128, 128), (128, 256, 64), (256, 256, 64), (256, 512, 32), (512, 512, 32),
          (512, 512, 16)]
tot = 0.0
for cin, cout, r in layers:
    w = torch.randn(cout, cin, 3, 3, device='cuda') * 0.05
    b = torch.zeros(cout, device='cuda')
    w.requires_grad_(False)
    p, tcf, tcd = _prep(w, b, 1.0, False, algo)
    res = []
    for name, nb, slab, ci, co, relu, bias in (('fwd', 2 * N, p.w_fwd, cin, cout, True, p.bias_p),
                                               ('dgrad', N, p.w_dgrad, cout, cin, False, None)):
        x = torch.randn(nb, r, r, ci, device='cuda')
        y = _empty(nb, r, r, co, 'cuda')
        st = stream_of(x)
        for _ in range(2):
            _vconv(st, x, slab, bias, y, nb, r, r, ci, co, relu, True, 'vgg')
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            _vconv(st, x, slab, bias, y, nb, r, r, ci, co, relu, True, 'vgg')
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 5 * 1e3
        tf = 2.0 * nb * r * r * ci * co * 9 / us / 1e6
        res.append(f'{name} {us:7.1f} us {tf:6.1f} TF/s')
        tot += us
    print(f'{cin:4d}->{cout:4d} @{r:3d}: ' + ' | '.join(res), flush=True)
print(f'sum of distinct shapes: {tot / 1e3:.2f} ms')
