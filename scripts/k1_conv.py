"""Development aid: the discriminator's 1x1 skip convolutions (forward with the residual add, transposed without) -- timing, and
the launch to point ncu at (`ncu --set full -k regex:conv_tc -c 2 python scripts/k1_conv.py once`)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
import torch
from b200gan._lib import lib, check, ptr, conv_workspace
from b200gan.modconv import _weight_prep

once = len(sys.argv) > 1 and sys.argv[1] == 'once'
dev = 'cuda'
B = 16
flush = torch.empty(64 << 20, device=dev)
shapes = [(128, 256, 128, True), (128, 256, 128, False), (256, 512, 64, True), (256, 512, 64, False), (256, 128, 128, False), (256, 128, 128, True),
          (512, 512, 32, True)]
if once:
    shapes = shapes[:1] + shapes[2:3]
for cin, cout, h, with_res in shapes:
    x = torch.randn(B, h, h, cin, device=dev)
    wt = torch.randn(1, cout, cin, 1, 1, device=dev)
    prep = _weight_prep(wt, None, 1.0 / cin ** 0.5, False, True, True, cin, cout, False)
    res = torch.randn(B, h, h, cout, device=dev) if with_res else None
    out = torch.empty(B, h, h, cout, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: check(lib.cagc_conv2d(st, x.data_ptr(), prep.w_fwd.data_ptr(), None, ptr(res), out.data_ptr(), B, h, h, cin, cout,
                                         cout, 1, 0, 0, 1.0, 1))
    call()
    torch.cuda.synchronize()
    if once:
        continue
    ref = torch.einsum('bhwc,oc->bhwo', x, wt.reshape(cout, cin)) / cin ** 0.5 + (res if with_res else 0)
    err = float((out - ref).abs().max() / ref.abs().max())
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    nbytes = 4.0 * B * h * h * (cin + cout * (2 if with_res else 1))
    print(f'{cin}->{cout} @{h} res={with_res}: {tot * 100:6.1f} us  {nbytes / (tot * 100) / 1e3:5.0f} GB/s  rel err {err:.1e}', flush=True)
