"""Timing of the face-parser branch of the KD step alone (development aid): preprocess -> FaceParser -> mask."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
import torch
from b200gan import maskglue
from b200gan.parsing import FaceParser, synthetic_state_dict

torch.backends.cudnn.benchmark = True
fp = FaceParser.from_state_dict(synthetic_state_dict(0)).cuda()
img = torch.tanh(torch.randn(16, 3, 256, 256, device='cuda'))


class Plain:
    def __init__(self, fp):
        self.fp = fp

    def __call__(self, x):
        return self.fp(x)


for name, fused, net in (('plain conv/add/relu + 512^2 scores', False, Plain(fp)), ('fused cuDNN + 512^2 scores', True, Plain(fp)),
                         ('fused cuDNN + low-res mask kernel', True, fp)):
    fp.fused = fused
    for _ in range(3):
        maskglue.content_mask(img, net)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        maskglue.content_mask(img, net)
    e1.record()
    torch.cuda.synchronize()
    print(f'{name}: {e0.elapsed_time(e1) / 10:.3f} ms per batch of 16', flush=True)
