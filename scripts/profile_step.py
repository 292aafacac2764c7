"""One KD step between cudaProfilerStart/Stop (ncu --profile-from-start off) after a few warm-up steps."""
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

size, B = 256, int(os.environ.get('BATCH', 16))
algo = config.best_available_algo() if os.environ.get('ALGO', 'tc') == 'tc' else config.ALGO_SIMT_FP32
config.set_default_algo(algo)
torch.manual_seed(0)
dev = torch.device('cuda')
teacher = model.Generator(size, 512, 8).to(dev)
student = model.Generator(size, 512, 8, generator_net_shape=bench.STUDENT_SHAPES[size]).to(dev)
disc = model.Discriminator(size).to(dev)
if os.environ.get('LOSS', 'full') == 'full':
    percept, parser = bench.kd_loss_networks(torch, dev)
    kd = KDStep(student, teacher, disc, percept_loss=percept, parsing_net=parser)
else:
    kd = KDStep(student, teacher, disc, mask=bench.synthetic_mask(size, dev))
kd.teacher_stream = None          # one stream: the launch list is in program order
z = lambda: [torch.randn(B, 512, device=dev), torch.randn(B, 512, device=dev)]
for _ in range(3):
    kd.step(z(), 5)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
kd.step(z(), 5)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
