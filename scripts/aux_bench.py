"""Development aid: the bandwidth-class kernels around the convolutions (epilogue backward, modulation backward, ToRGB
forward / backward, modulate) on the KD-step shapes, one launch at a time with an L2 flush in between."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
import torch
from b200gan._lib import lib, check, ptr

dev = 'cuda'
B = 16
flush = torch.empty(64 << 20, device=dev)
fir = (torch.tensor([1., 3., 3., 1.])[:, None] * torch.tensor([1., 3., 3., 1.])[None, :] / 64 * 4).to(dev)


def timed(call, nbytes, reps=8):
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = tot / reps * 1e3
    return f'{us:6.1f}us {nbytes / us / 1e3:5.0f}GB/s'


shapes = [(256, 39), (128, 77), (64, 154), (32, 154), (256, 128), (128, 256), (64, 512)]
only = sys.argv[1:] or None
for h, c in shapes:
    p = (c + 7) // 8 * 8
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(B, h, h, p, device=dev)
    g = torch.randn(B, h, h, p, device=dev)
    o = torch.empty_like(x)
    s = torch.rand(B, p, device=dev) + 0.5
    d = torch.rand(B, p, device=dev) + 0.5
    noise = torch.randn(B, 1, h, h, device=dev)
    nw = torch.randn(1, device=dev)
    bias = torch.randn(p, device=dev)
    n = 4.0 * B * h * h * p
    line = [f'{h:3d}^2 x {c:3d}']
    chunks = lib.cagc_act_bwd_chunks(h, h)
    part = torch.empty(B, chunks, 3, p, device=dev)
    line.append('act_bwd ' + timed(lambda: check(lib.cagc_act_bwd(st, g.data_ptr(), h * h * p, 1, h * p, p, x.data_ptr(), d.data_ptr(),
                                                                  noise.data_ptr(), nw.data_ptr(), bias.data_ptr(), o.data_ptr(),
                                                                  part.data_ptr(), B, h, h, p, c, h * h, 1)), 3 * n))
    mpart = torch.empty(B, chunks, p, device=dev)
    line.append('mod_bwd ' + timed(lambda: check(lib.cagc_mod_bwd(st, g.data_ptr(), x.data_ptr(), s.data_ptr(), mpart.data_ptr(),
                                                                  B, h, h, p)), 3 * n))
    line.append('modulate ' + timed(lambda: check(lib.cagc_modulate(st, x.data_ptr(), s.data_ptr(), o.data_ptr(), B, h, h, p)), 2 * n))
    w2 = torch.randn(3, c, device=dev)
    rb = torch.randn(3, device=dev)
    skip = torch.randn(B, 3, h // 2, h // 2, device=dev)
    rgb = torch.empty(B, 3, h, h, device=dev)
    line.append('torgb_fwd ' + timed(lambda: check(lib.cagc_torgb_fwd(st, x.data_ptr(), w2.data_ptr(), s.data_ptr(), rb.data_ptr(),
                                                                      skip.data_ptr(), fir.data_ptr(), rgb.data_ptr(), B, h, h, p, c, 3,
                                                                      0.1, 4, 4, 2, 1)), n))
    grgb = torch.randn(B, 3, h, h, device=dev)
    tpart = torch.empty(B, chunks, 3, p, device=dev)
    line.append('torgb_bwd ' + timed(lambda: check(lib.cagc_torgb_bwd(st, grgb.data_ptr(), x.data_ptr(), w2.data_ptr(), s.data_ptr(),
                                                                      g.data_ptr(), o.data_ptr(), tpart.data_ptr(), B, h, h, p, c, 3,
                                                                      0.1)), 3 * n))
    print(' | '.join(line), flush=True)
