"""Development aid: NHWC FIR kernel variants (env knobs of cagc_tc_fir_nhwc) on the KD-step shapes."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
import torch
from b200gan._lib import lib, check, fir_nhwc

dev = 'cuda'
B = 16
shapes = [(129, 256, 1), (257, 128, 1), (257, 40, 1), (129, 80, 1), (65, 160, 1), (256, 128, 2), (128, 256, 2), (64, 512, 2)]
fir = (torch.tensor([1., 3., 3., 1.])[:, None] * torch.tensor([1., 3., 3., 1.])[None, :] / 64).to(dev)
configs = [dict(CAGC_FIR_IMPL='stream', CAGC_FIR_CW='32'), dict(CAGC_FIR_IMPL='stream', CAGC_FIR_CW='64'),
           dict(CAGC_FIR_IMPL='stream', CAGC_FIR_CW='128'),
           dict(CAGC_FIR_IMPL='stream', CAGC_FIR_CW='32', CAGC_FIR_CTAS='4736'),
           dict(CAGC_FIR_IMPL='stream', CAGC_FIR_CW='32', CAGC_FIR_CTAS='592'),
           dict(CAGC_FIR_IMPL='stream', CAGC_FIR_CW='32', CAGC_FIR_STAGES='4'),
           dict(CAGC_FIR_IMPL='ldg')]
knobs = ['CAGC_FIR_IMPL', 'CAGC_FIR_CW', 'CAGC_FIR_CTAS', 'CAGC_FIR_STAGES']
bufs = {}
for h, p, pad in shapes:
    ho = h + 2 * pad - 3
    bufs[(h, p, pad)] = (torch.randn(B, h, h, p, device=dev), torch.empty(B, ho, ho, p, device=dev), ho)
ref = {}
for cfg in configs:
    for k in knobs:
        os.environ.pop(k, None)
    os.environ.update(cfg)
    line = []
    for key in shapes:
        h, p, pad = key
        x, y, ho = bufs[key]
        st = torch.cuda.current_stream().cuda_stream
        call = lambda: fir_nhwc(st, x.data_ptr(), fir, None, None, None, None, y.data_ptr(), B, h, h, p, p,
                                (pad, pad, pad, pad), 0, 0, 'fir')
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        if key not in ref:
            ref[key] = y.clone()
        else:
            assert torch.equal(ref[key], y), (cfg, key)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        gbs = 4.0 * B * p * (h * h + ho * ho) / us / 1e3
        line.append(f'{us:6.0f}us {gbs:5.0f}')
    print(' '.join(f'{k[9:]}={v}' for k, v in cfg.items()).ljust(40), ' | '.join(line), flush=True)
print('shapes (h, pitch, pad):', shapes)
