"""One LPIPS-VGG16 forward + backward (batch 16, 256px) and one content-mask pass between cudaProfilerStart/Stop
(ncu --profile-from-start off) after warm-up."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
sys.path.insert(0, ROOT)
import torch
import bench
from b200gan import config, maskglue

config.set_default_algo(config.best_available_algo())
dev = torch.device('cuda')
percept, parser = bench.kd_loss_networks(torch, dev)
B = int(os.environ.get('BATCH', 16))
t_img = torch.tanh(torch.randn(B, 3, 256, 256, device=dev))
s_img = (t_img + 0.2 * torch.randn_like(t_img)).clamp(-1, 1).requires_grad_(True)


def once():
    mask = maskglue.content_mask(t_img, parser)
    loss = 3.0 * torch.mean(percept(s_img * mask, t_img * mask))
    loss.backward()


for _ in range(3):
    once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
