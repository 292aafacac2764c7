"""Development aid: a few launches of the NHWC FIR on one KD-step shape (for `ncu --set full`)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
import torch
from b200gan._lib import fir_nhwc

dev = 'cuda'
B, h, p, pad = 16, int(os.environ.get('FIR_H', 257)), int(os.environ.get('FIR_P', 128)), 1
fir = (torch.tensor([1., 3., 3., 1.])[:, None] * torch.tensor([1., 3., 3., 1.])[None, :] / 64).to(dev)
ho = h + 2 * pad - 3
x = torch.randn(B, h, h, p, device=dev)
y = torch.empty(B, ho, ho, p, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    fir_nhwc(st, x.data_ptr(), fir, None, None, None, None, y.data_ptr(), B, h, h, p, p, (pad, pad, pad, pad), 0, 0, 'fir')
torch.cuda.synchronize()
print('done')
