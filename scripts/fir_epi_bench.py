"""Development aid: cost of the FIR epilogue terms (demod scale / noise / bias / activation) on the Blur shapes of the KD step."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
import torch
from b200gan._lib import fir_nhwc

dev = 'cuda'
B = 16
shapes = [(257, 128, 1), (129, 256, 1), (257, 40, 1), (129, 80, 1)]
fir = (torch.tensor([1., 3., 3., 1.])[:, None] * torch.tensor([1., 3., 3., 1.])[None, :] / 64).to(dev)
flush = torch.empty(64 << 20, device=dev)
variants = ['plain', 'scale', 'bias', 'act', 'bias+act', 'noise', 'all']
for h, p, pad in shapes:
    ho = h + 2 * pad - 3
    x = torch.randn(B, h, h, p, device=dev)
    y = torch.empty(B, ho, ho, p, device=dev)
    d = torch.rand(B, p, device=dev) + 0.5
    noise = torch.randn(B, 1, ho, ho, device=dev)
    nw = torch.randn(1, device=dev)
    bias = torch.randn(p, device=dev)
    line = []
    for v in variants:
        a = v == 'all'
        args = (d if (a or v == 'scale') else None, noise if (a or v == 'noise') else None, nw if (a or v == 'noise') else None,
                bias if (a or 'bias' in v) else None)
        act = 1 if (a or 'act' in v) else 0
        st = torch.cuda.current_stream().cuda_stream
        call = lambda: fir_nhwc(st, x.data_ptr(), fir, args[0], args[1], args[2], args[3], y.data_ptr(), B, h, h, p, p,
                                (pad, pad, pad, pad), ho * ho, act, 'fir')
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        us = tot * 100
        gbs = 4.0 * B * p * (h * h + ho * ho) / us / 1e3
        line.append(f'{v}: {us:5.0f}us {gbs:5.0f}')
    print((h, p, pad), ' | '.join(line), flush=True)
