"""Development aid: per-parameter gradient agreement of the tcgen05 path with the exact-fp32 path at 256px."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
import torch
import model
from b200gan import config
torch.manual_seed(11)
shape = [154] * 10 + [77, 77, 39, 39]
b = 4
gen = model.Generator(256, 512, 8, generator_net_shape=shape).cuda()
gsd = torch.Generator().manual_seed(5)
with torch.no_grad():
    for n, p in gen.named_parameters():
        if n.endswith('noise.weight') or n.endswith('activate.bias') or (n.endswith('.bias') and p.ndim == 4):
            p.copy_(torch.randn(p.shape, generator=gsd) * 0.3)
z = [torch.randn(b, 512, device='cuda'), torch.randn(b, 512, device='cuda')]
noise = [torch.randn(b, 1, n.shape[2], n.shape[3], device='cuda') for n in gen.make_noise()]
cot = torch.randn(b, 3, 256, 256, device='cuda')
res = {}
for name, algo in (('fp32', 0), ('fp32b', 0), ('tc', 1)):
    gen.zero_grad()
    with config.use_algo(algo):
        out = gen(z, inject_index=5, noise=noise, return_rgb_list=True)
        ((out[-1] * cot).abs().mean() * 3 + sum(r.mean() for r in out[:-1])).backward()
    res[name] = {n: p.grad.detach().clone() for n, p in gen.named_parameters()}
for n, g in res['fp32'].items():
    l2 = float((res['tc'][n] - g).norm() / g.norm().clamp_min(1e-20))
    l2b = float((res['fp32b'][n] - g).norm() / g.norm().clamp_min(1e-20))
    flag = '  <<<' if l2 > 3e-2 else ''
    if l2 > 1e-2 or 'noise' in n:
        print(f'{n:40s} numel {g.numel():8d} |g| {float(g.norm()):.3e}  tc-vs-fp32 {l2:.3e}  fp32 rerun {l2b:.1e}{flag}')
