#!/usr/bin/env python
"""The benchmarked KD-like generator step executed by the UNMODIFIED reference modules staged under
oracle/_ref/ (oracle/stage_ref.py) -- TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/ref_step.py --device cpu  --batch 2  --steps 2  --warmup 1     # reference CPU path
    python oracle/ref_step.py --device cuda --batch 16 --steps 50 --warmup 10    # reference CUDA path on the B200

The step is train.py:280-308 composed from the reference's own `model.Generator` / `model.Discriminator`
(its `op/` kernels, cuDNN grouped convolutions / native CPU fallbacks) exactly as bench.py's product arm composes
it: student forward with rgb list -> discriminator -> g_nonsaturating_loss (train.py:215-218); teacher forward
(eval, requires_grad False: train.py:500-503); then, with `--loss full` (default), the reference's OWN `KD_loss`
(train.py:145-184, its function text executed from the staged train.py) with its own `lpips.PerceptualLoss`
(net-lin VGG16: torchvision architecture, random weights -- the pretrained ones cannot be downloaded -- and the vendored
lin heads), its own BiSeNet class (random weights: 79999_iter.pth is not staged), `Batch_Img_Parsing` and
`Get_Masked_Tensor`; with `--loss kdlike` a constant content mask and the L1 term only (rounds 1-2 workload);
generator.zero_grad(); backward; torch.optim.Adam with the reference's hyper-parameters (train.py:528-532).

Runs in its own process: `model` / `op` here are the REFERENCE's modules (the product's drop-in modules have the
same names).  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')

STUDENT_SHAPES = {256: [154] * 10 + [77, 77, 39, 39],
                  1024: [154] * 10 + [77, 77, 39, 39, 20, 20, 10, 10]}


def import_reference(device):
    """Import the staged reference `model`.  Its op/ modules JIT-load two CUDA extensions at import; they were
    pre-built by stage_ref.py into oracle/_ref/_ext (rebuilt here by torch if the cache does not match)."""
    if not os.path.exists(os.path.join(REF, 'model.py')):
        raise RuntimeError('reference not staged: run `python oracle/stage_ref.py` in the build container')
    os.environ['TORCH_EXTENSIONS_DIR'] = os.path.join(REF, '_ext')
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    for name in ('model', 'op'):
        assert name not in sys.modules, f'{name} already imported: the reference needs its own process'
    sys.path.insert(0, REF)
    how = 'jit extensions (prebuilt cache)'
    try:
        import model as ref_model
    except Exception as e:                       # no nvcc / stale cache on a CPU-only run: the CPU path never calls the ops
        if device != 'cpu':
            raise
        for k in [k for k in sys.modules if k == 'op' or k.startswith('op.') or k == 'model']:
            del sys.modules[k]
        import torch.utils.cpp_extension as ce
        ce.load = lambda *a, **k: None
        import model as ref_model
        how = f'extension build skipped ({type(e).__name__}); CPU fallbacks do not call it'
    assert os.path.realpath(ref_model.__file__).startswith(os.path.realpath(REF))
    return ref_model, how


def reference_kd_loss(device_str):
    """(KD_loss, percept_loss, parsing_net) built from the reference's own code.  train.py parses the command line and
    builds datasets at import, so its function definitions are executed from the source text instead."""
    import ast
    import types
    import torch
    import torch.nn.functional as F
    import numpy as np
    for name in ('skimage', 'skimage.measure', 'skimage.color', 'skimage.transform', 'IPython'):
        if name not in sys.modules:                  # lpips/__init__.py:7, networks_basic.py:11-12 (unused by net-lin)
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.__path__ = []
                sys.modules[name] = m
    sys.modules['skimage.measure'].__dict__.setdefault('compare_ssim', None)
    sys.modules['IPython'].__dict__.setdefault('embed', lambda *a, **k: None)
    import torchvision.models as tvm
    real_vgg = tvm.vgg16
    tvm.vgg16 = lambda pretrained=True, **k: real_vgg(weights=None)        # no network: random VGG16
    import torch.utils.model_zoo as mz
    mz.load_url = lambda *a, **k: {}                                        # resnet.py:83: random ResNet18
    import lpips
    import train_hyperparams
    from Util.content_aware_pruning import Batch_Img_Parsing, Get_Masked_Tensor
    from Util.face_parsing.BiSeNet import BiSeNet
    ns = {'torch': torch, 'F': F, 'np': np, 'train_hyperparams': train_hyperparams, 'device': device_str,
          'Batch_Img_Parsing': Batch_Img_Parsing, 'Get_Masked_Tensor': Get_Masked_Tensor}
    path = os.path.join(REF, 'train.py')
    for node in ast.parse(open(path).read()).body:
        if isinstance(node, ast.FunctionDef) and node.name in ('KD_loss', 'Downsample_Image_256'):
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), ns)
    gpu = device_str != 'cpu'
    percept = lpips.PerceptualLoss(model='net-lin', net='vgg', use_gpu=gpu, gpu_ids=[0])       # train.py:510
    parsing_net = BiSeNet(n_classes=19).to(device_str).eval()                                   # content_aware_pruning.py:24-28
    return ns['KD_loss'], percept, parsing_net


def synthetic_mask(size, device):
    import torch
    yy, xx = torch.meshgrid(torch.arange(size, device=device), torch.arange(size, device=device), indexing='ij')
    c = (size - 1) / 2
    return ((((yy - c) / (0.42 * size)) ** 2 + ((xx - c) / (0.34 * size)) ** 2) <= 1).float().view(1, 1, size, size)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--device', default='cpu', choices=['cpu', 'cuda'])
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--batch', type=int, default=2)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--threads', type=int, default=0)
    ap.add_argument('--tf32', type=int, default=1, help='cuDNN / cuBLAS TF32 (torch default for convolutions: on)')
    ap.add_argument('--loss', default='full', choices=['full', 'kdlike'])
    args = ap.parse_args()

    import torch
    import torch.nn.functional as F
    cores = args.threads or (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    ref_model, how = import_reference(args.device)
    dev = torch.device(args.device)
    if args.device == 'cuda':
        torch.backends.cudnn.allow_tf32 = bool(args.tf32)
        torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
        torch.backends.cudnn.benchmark = True
    torch.manual_seed(1234)
    size, B = args.size, args.batch
    teacher = ref_model.Generator(size, 512, 8).to(dev).eval()
    student = ref_model.Generator(size, 512, 8, generator_net_shape=STUDENT_SHAPES[size]).to(dev)
    disc = ref_model.Discriminator(size).to(dev)
    for p in teacher.parameters():
        p.requires_grad_(False)                                  # train.py:502-503
    g_optim = torch.optim.Adam(student.parameters(), lr=0.002 * 0.8, betas=(0.0, 0.99 ** 0.8))   # train.py:528-532
    mask = synthetic_mask(size, dev)
    inject = 5
    full = args.loss == 'full'
    if full:
        kd_loss_fn, percept, parsing_net = reference_kd_loss('cuda:0' if args.device == 'cuda' else 'cpu')
        kd_args = argparse.Namespace(kd_mode='Output_Only', kd_l1_lambda=3.0, kd_lpips_lambda=3.0, size=size)

    def step(z):
        for p in student.parameters():
            p.requires_grad_(True)                               # train.py:286
        for p in disc.parameters():
            p.requires_grad_(False)                              # train.py:287
        fake_list = student(z, return_rgb_list=True, inject_index=inject)          # train.py:291
        g_loss = F.softplus(-disc(fake_list[-1])).mean()                          # train.py:293-294
        if full:
            l1, lp = kd_loss_fn(kd_args, teacher, z, inject, fake_list[-1], fake_list, percept, parsing_net)   # train.py:301
            total = g_loss + l1 + lp                                                # train.py:304
        else:
            real = teacher(z, return_rgb_list=True, inject_index=inject)[-1]      # train.py:151-152
            kd = 3.0 * torch.mean(torch.abs(real * mask - fake_list[-1] * mask))  # train.py:157-164
            total = g_loss + kd
        student.zero_grad()                                                       # train.py:306
        total.backward()
        g_optim.step()
        return total

    def sync():
        if args.device == 'cuda':
            torch.cuda.synchronize()

    def latents():
        return [torch.randn(B, 512, device=dev), torch.randn(B, 512, device=dev)]

    for _ in range(args.warmup):
        step(latents())
    sync()
    zs = [latents() for _ in range(args.steps)]
    sync()
    if args.device == 'cuda':
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.perf_counter()
    for z in zs:
        loss = step(z)
    if args.device == 'cuda':
        e1.record()
        sync()
        sec = e0.elapsed_time(e1) / 1e3
    else:
        sec = time.perf_counter() - t0
    out = {'impl': 'reference', 'device': args.device, 'size': size, 'batch': B, 'steps': args.steps,
           'warmup': args.warmup, 'sec_per_step': sec / args.steps, 'images_per_s': B * args.steps / sec,
           'cores': cores, 'loss': float(loss.detach()), 'loss_kind': args.loss, 'import': how, 'torch': torch.__version__,
           'tf32': bool(args.tf32) if args.device == 'cuda' else None,
           'source': 'oracle/_ref (byte-identical copy of the reference, see MANIFEST.json)'}
    if args.device == 'cuda':
        out['gpu'] = torch.cuda.get_device_name(0)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
