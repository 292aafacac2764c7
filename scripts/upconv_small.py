"""Development aid: low-resolution transposed convolutions (phased split-K path) -- parity against the exact-fp32 engine and timing."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
import torch
from b200gan._lib import lib, check, ptr, conv_workspace
from b200gan.modconv import _weight_prep

dev = 'cuda'
flush = torch.empty(64 << 20, device=dev)
for B in (16, 2):
    for cin, cout, h in [(512, 512, 4), (512, 512, 8), (512, 512, 16), (154, 154, 4), (154, 154, 8), (154, 154, 16), (154, 154, 32),
                         (512, 512, 32), (520, 264, 5)]:
        pin, pout = (cin + 7) // 8 * 8, (cout + 7) // 8 * 8
        g = torch.Generator().manual_seed(cin + h)
        x = torch.zeros(B, h, h, pin, device=dev)
        x[..., :cin] = torch.randn(B, h, h, cin, generator=g).to(dev)
        wt = torch.randn(1, cout, cin, 3, 3, generator=g).to(dev)
        outs = {}
        for algo in (0, 1):
            tc = bool(algo)
            prep = _weight_prep(wt, None, 1.0 / (cin * 9) ** 0.5, True, tc, tc, pin, pout, False)
            hu = 2 * h + 1
            ut = torch.full((B, hu, hu, pout), float('nan'), device=dev)
            st = torch.cuda.current_stream().cuda_stream
            ws, wsb = conv_workspace(B, hu, hu, pout, dev) if tc else (None, 0)
            ones = torch.ones(B, pin, device=dev)
            call = lambda: check(lib.cagc_conv_up_ws(st, x.data_ptr(), prep.w_fwd.data_ptr(), None if tc else ones.data_ptr(),
                                                     ut.data_ptr(), B, h, h, pin, pout, 3, algo, ptr(ws), wsb))
            call()
            torch.cuda.synchronize()
            outs[algo] = ut.clone()
            if algo == 1:
                tot = 0.0
                for _ in range(10):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    call()
                    e1.record()
                    torch.cuda.synchronize()
                    tot += e0.elapsed_time(e1)
                os.environ['X'] = '1'
        err = float((outs[1] - outs[0]).abs().max() / outs[0].abs().max())
        print(f'B={B:2d} {cin:3d}->{cout:3d} @{h:2d}: {tot * 100:6.1f} us  rel err {err:.2e}  ws={wsb >> 20} MB', flush=True)
        assert err < 3e-3 and torch.isfinite(outs[1]).all()
