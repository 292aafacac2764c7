"""Development aid: kernel census of one KD step (torch.profiler, eager launches) next to the graph-replay time."""
import collections
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
sys.path.insert(0, ROOT)
import torch
import bench
import model
from b200gan import config
from b200gan.kd import KDStep
from torch.profiler import profile, ProfilerActivity

size, B = 256, int(os.environ.get('BATCH', 16))
config.set_default_algo(config.best_available_algo())
torch.manual_seed(0)
dev = torch.device('cuda')
teacher = model.Generator(size, 512, 8).to(dev)
student = model.Generator(size, 512, 8, generator_net_shape=bench.STUDENT_SHAPES[size]).to(dev)
disc = model.Discriminator(size).to(dev)
kd = KDStep(student, teacher, disc, mask=bench.synthetic_mask(size, dev))
z = lambda: [torch.randn(B, 512, device=dev), torch.randn(B, 512, device=dev)]
for _ in range(3):
    kd.step(z(), 5)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    kd.step(z(), 5)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name[:90]][0] += 1
        agg[e.name[:90]][1] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
tot_n = sum(v[0] for v in agg.values())
tot_t = sum(v[1] for v in agg.values())
print(f'kernels/step {tot_n}, summed kernel time {tot_t / 1e3:.2f} ms')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get('TOP', 60))]:
    print(f'{n:5d} {t:9.1f} us  {k}')
kd.capture(B, 5)
for _ in range(3):
    kd.step_graphed(z())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    kd.step_graphed(z())
e1.record()
torch.cuda.synchronize()
print(f'graph replay {e0.elapsed_time(e1) / 10:.2f} ms/step')
