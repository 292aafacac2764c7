#!/usr/bin/env python
"""Benchmark of the KD retraining generator step (BASELINE.json metric: images/sec, 256px 70%-pruned
StyleGAN2 student, full-size teacher, global batch 16) -- see DESIGN.md "Measurement".

    python bench.py --gpus N --steps K --warmup W                 # product arm (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...       # the reference's OWN code on the host CPU cores
    python bench.py --config kd1024|fid256|saliency256 ...        # BASELINE.json configs[3], [4], [2]

One step = train.py:280-308 (BASELINE.json configs[1]): student forward (RGB list) + discriminator forward + GAN loss,
teacher forward, BiSeNet face parsing of the teacher image at 512px + content mask, masked L1 + LPIPS-VGG16 KD losses,
backward, gradient all-reduce over ranks, fused Adam.  The pretrained VGG16 / BiSeNet weights are not available offline:
both networks run with random weights of the same architecture on every arm.  `--loss kdlike` is the workload of rounds
1-2 (no LPIPS, no parser; constant content mask); the default run reports it as well under "kd_like".

Scaling: the north star shards the minibatch of 16 over the GPUs of one box, so the headline mode is STRONG
scaling (global batch 16, per-rank 16/N); the weak-scaling figure (16 per GPU) is measured in the same run and
reported under "weak_scaling".  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'content-aware-gan-compression_b200')

if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
    del os.environ['NCCL_DEBUG']           # both print NCCL's version banner on stdout: rank 0 prints ONE JSON line

STUDENT_SHAPES = {256: [154] * 10 + [77, 77, 39, 39],
                  1024: [154] * 10 + [77, 77, 39, 39, 20, 20, 10, 10]}
# conv FLOPs per image, reference convention (Util/Calculators.py:16-61 x2), SURVEY.md §8d
GFLOP_PER_IMG = {256: {'student': 8.240, 'teacher': 90.236}, 1024: {'student': 13.973, 'teacher': 148.520}}
GLOBAL_BATCH = 16
REF_STEP = os.path.join(ROOT, 'oracle', 'ref_step.py')


def kd_workload(size, loss='full'):
    s = STUDENT_SHAPES[size]
    if loss == 'full':
        return (f'KD generator step {size}px (train.py:280-308): 70%-pruned student {s[0]}/{s[-1]}ch f+b, full teacher f, '
                f'discriminator f+dgrad, BiSeNet parsing @512 + content mask, masked L1 + LPIPS-VGG16'
                f'{" on bilinear-256" if size > 256 else ""} KD losses (random-init VGG16 / BiSeNet weights), '
                f'grad all-reduce, Adam; global batch {GLOBAL_BATCH}; style mixing inject_index=5')
    return (f'KD-like generator step {size}px: 70%-pruned student {s[0]}/{s[-1]}ch f+b, full teacher f, '
            f'discriminator f+dgrad, masked L1 KD, grad all-reduce, Adam; global batch {GLOBAL_BATCH}; '
            f'style mixing inject_index=5')


def kd_config(size, world, loss='full'):
    """`config` of the JSON line -- identical on the product arm and on the reference arm."""
    return {'workload': kd_workload(size, loss), 'global_batch': GLOBAL_BATCH, 'parallelism': f'dp{world}',
            'l2': 'per-step working set (>=4 GB of activations) exceeds the 126 MB L2; no explicit flush'}


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'],
                'bf16_tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arms: the reference's own code (oracle/_ref, staged by oracle/stage_ref.py) in its own process
# ------------------------------------------------------------------------------------------------
def run_ref_step(device, size, batch, steps, warmup, timeout=1500, loss='full'):
    """oracle/ref_step.py in a subprocess (the reference's `model` / `op` module names collide with the
    drop-in's).  Returns the parsed JSON or {'unavailable': why}."""
    cmd = [sys.executable, REF_STEP, '--device', device, '--size', str(size), '--batch', str(batch),
           '--steps', str(steps), '--warmup', str(warmup), '--loss', loss]
    env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MASTER_ADDR',
                                                           'MASTER_PORT', 'CAGC_CONV_ALGO')}
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, env=env)
    except subprocess.TimeoutExpired:
        return {'unavailable': f'{" ".join(cmd[1:])} timed out after {timeout} s'}
    for ln in reversed(r.stdout.strip().splitlines()):
        if ln.startswith('{'):
            return json.loads(ln)
    tail = (r.stderr or r.stdout or '').strip().splitlines()[-1:] or ['no output']
    return {'unavailable': f'rc {r.returncode}: {tail[0][:300]}'}


def cpu_baseline_from(ref, sample_b, steps, warmup):
    if 'unavailable' in ref:
        return None
    what = ("its own KD_loss / lpips / BiSeNet / Batch_Img_Parsing / Get_Masked_Tensor, same KD step"
            if ref.get('loss_kind') == 'full' else 'same KD-like step')
    return {'value': ref['images_per_s'], 'unit': 'images/s', 'cores': ref['cores'], 'kind': 'reference',
            'sample': f"the reference's own model.py + op/ CPU fallbacks (oracle/_ref, unmodified; {ref['import']}), "
                      f"{what}, bounded sample: batch {sample_b} per step, {warmup} warm-up + {steps} "
                      f"timed steps, torch {ref['torch']} CPU fp32, {ref['cores']} threads, "
                      f"{ref['sec_per_step']:.2f} s/step"}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the step, all host threads.  This process
    imports neither the product package nor the oracle restatement."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample_b = 2            # >= 2: Util/content_aware_pruning.py:108 squeezes the batch dim at 1
    warm = max(1, min(args.warmup, 2))
    ref = run_ref_step('cpu', args.size, sample_b, args.steps, warm, loss=args.loss)
    if 'unavailable' in ref:
        print(json.dumps({'impl': 'reference', 'unavailable': ref['unavailable']}))
        return
    ips, sec = ref['images_per_s'], ref['sec_per_step']
    line = {
        'impl': 'reference', 'metric': f'images/sec KD step {args.size}px pruned StyleGAN2', 'value': ips,
        'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': kd_config(args.size, args.gpus, args.loss),
        'cpu_baseline': cpu_baseline_from(ref, sample_b, args.steps, warm),
        'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def synthetic_mask(size, device):
    """Stand-in for the BiSeNet face mask (Util/content_aware_pruning.py:61-117): centred ellipse."""
    import torch
    yy, xx = torch.meshgrid(torch.arange(size, device=device), torch.arange(size, device=device), indexing='ij')
    c = (size - 1) / 2
    return ((((yy - c) / (0.42 * size)) ** 2 + ((xx - c) / (0.34 * size)) ** 2) <= 1).float().view(1, 1, size, size)


def kd_loss_networks(torch, dev):
    """LPIPS-VGG16 on the package's own kernels and the BiSeNet-architecture face parser (library convolutions), both
    with random weights of the reference architectures (torchvision's VGG16 initialiser: kaiming-normal fan-out, zero
    bias; non-negative lin heads as in lpips/weights/v0.1/vgg.pth).  Returns (percept_loss, parsing_net)."""
    from b200gan.lpips import PerceptualLossVGG, VGG16_CFG
    from b200gan.parsing import FaceParser, synthetic_state_dict
    g = torch.Generator().manual_seed(99)
    cws, cbs, cin = [], [], 3
    for c in [c for c in VGG16_CFG if c != 'M']:
        cws.append(torch.randn(c, cin, 3, 3, generator=g) * (2.0 / (9 * c)) ** 0.5)
        cbs.append(torch.zeros(c))
        cin = c
    lins = [torch.rand(c, generator=g) for c in (64, 128, 256, 512, 512)]
    percept = PerceptualLossVGG(cws, cbs, lins).to(dev)
    parser = FaceParser.from_state_dict(synthetic_state_dict(0)).to(dev)
    torch.backends.cudnn.benchmark = True           # the parser's library convolutions (autotuned during warm-up)
    return percept, parser


class Ctx:
    """Per-process setup shared by the product-arm configs."""

    def __init__(self, args):
        for p in (PKG, ROOT):
            if p not in sys.path:
                sys.path.insert(0, p)
        import torch
        from b200gan import dist as D, config, _lib
        self.torch, self.D, self.config, self.lib = torch, D, config, _lib
        self.local = D.init_from_env()
        self.rank, self.world = D.get_rank(), D.get_world_size()
        assert torch.cuda.is_available(), 'bench.py (impl b200) needs a CUDA device; there is no CPU fallback'
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        algo = {'simt': config.ALGO_SIMT_FP32, 'tc': config.ALGO_TCGEN05_TF32}.get(args.algo)
        if algo is None:
            algo = config.best_available_algo()
        config.set_default_algo(algo)
        self.algo = algo
        self.tf32 = algo == config.ALGO_TCGEN05_TF32
        torch.manual_seed(1234 + self.rank)

    def barrier_sync(self):
        self.D.synchronize()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, ms):
        t = self.torch.tensor([ms], device=self.dev)
        if self.world > 1:
            self.torch.distributed.all_reduce(t, op=self.torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def time_steps(self, fn, n):
        """EXACTLY n calls of fn(i) between barrier + synchronize on both sides; device time, max over ranks."""
        torch = self.torch
        self.barrier_sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        self.barrier_sync()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def broadcast_module(self, m):
        if self.world > 1:
            for t in list(m.parameters()) + list(m.buffers()):
                self.torch.distributed.broadcast(t.data, 0)


def bench_kd(args, cx):
    torch, D, config, _lib = cx.torch, cx.D, cx.config, cx.lib
    from b200gan.kd import KDStep
    import model
    rank, world, dev = cx.rank, cx.world, cx.dev
    size = args.size
    strong = args.scaling == 'strong'
    if strong:
        assert GLOBAL_BATCH % world == 0, f'global batch {GLOBAL_BATCH} does not shard over {world} ranks'
        B = GLOBAL_BATCH // world
    else:
        B = GLOBAL_BATCH
    args.warmup = max(args.warmup, 3)
    inject = 5

    teacher = model.Generator(size, 512, 8).to(dev)
    student = model.Generator(size, 512, 8, generator_net_shape=STUDENT_SHAPES[size]).to(dev)
    disc = model.Discriminator(size).to(dev)
    for m in (teacher, student, disc):
        cx.broadcast_module(m)                      # identical replicas on every rank
    full = args.loss == 'full'
    percept, parser = kd_loss_networks(torch, dev) if full else (None, None)
    kd = KDStep(student, teacher, disc, percept_loss=percept, parsing_net=parser,
                mask=None if full else synthetic_mask(size, dev))

    def fresh_latents(b):
        return [torch.randn(b, 512, device=dev), torch.randn(b, 512, device=dev)]

    use_graph = not args.no_graph

    def measure(kd, B, steps, warmup, with_clocks):
        """value leg: device-resident latents.  Returns (ms_total, launches, clocks)."""
        if use_graph:
            kd.capture(B, inject)            # includes 3 eager warm-up steps on a side stream
            run_step = lambda z: kd.step_graphed(z)
        else:
            run_step = lambda z: kd.step(z, inject)
        for _ in range(warmup):
            run_step(fresh_latents(B))
        lat = [fresh_latents(B) for _ in range(min(steps, 64))]
        sampler = ClockSampler(cx.local) if (with_clocks and rank == 0) else None
        cx.barrier_sync()
        if sampler:
            sampler.start()
        n0 = _lib.launch_count()
        ms_total = cx.time_steps(lambda i: run_step(lat[i % len(lat)]), steps)
        launches = (kd.launches_per_replay * steps) if use_graph else (_lib.launch_count() - n0)
        return ms_total, launches, (sampler.stop() if sampler else None), run_step, lat

    ms_total, launches, clocks, run_step, lat = measure(kd, B, args.steps, args.warmup, True)
    value = world * B * args.steps / (ms_total / 1e3)

    # ---------------- end-to-end through the public API: pinned host latents in, loss out, every step
    hz = [[torch.randn(B, 512).pin_memory(), torch.randn(B, 512).pin_memory()] for _ in range(min(args.steps, 64))]
    for _ in range(2):
        kd.step_from_host(hz[0], inject)
    e2e_ms = cx.time_steps(lambda i: kd.step_from_host(hz[i % len(hz)], inject), args.steps)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)

    # ---------------- sustained leg: the same step for >= 200 iterations (clocks settle under the power cap)
    sustained = None
    if args.sustained_steps > 0:
        sampler = ClockSampler(cx.local) if rank == 0 else None
        cx.barrier_sync()
        if sampler:
            sampler.start()
        sus_ms = cx.time_steps(lambda i: run_step(lat[i % len(lat)]), args.sustained_steps)
        sc = sampler.stop() if sampler else None
        sustained = {'steps': args.sustained_steps, 'ms_per_step': sus_ms / args.sustained_steps,
                     'value': world * B * args.sustained_steps / (sus_ms / 1e3), 'unit': 'images/s',
                     'timed_region_s': sus_ms / 1e3, 'clocks': sc}

    # ---------------- per-kernel CUDA events (roofline): eager steps, events on the launching stream
    prof = config.KernelProfiler()
    overlap_stream, kd.teacher_stream = kd.teacher_stream, None   # one stream: a kernel's events bracket only itself
    kd.step(lat[0], inject)
    cx.barrier_sync()
    config.set_profiler(prof)
    prof_steps = min(args.steps, 5)
    for i in range(prof_steps):
        kd.step(lat[i % len(lat)], inject)
    cx.barrier_sync()
    config.set_profiler(None)
    kd.teacher_stream = overlap_stream

    # ---------------- generator slice alone (student f+b, teacher f) for the Amdahl picture
    def slice_step(z):
        kd.bucket.detach_grads()
        fake = student(z, return_rgb_list=True, inject_index=inject)
        with torch.no_grad():
            real = teacher(z, return_rgb_list=True, inject_index=inject)
        (3.0 * (real[-1] - fake[-1]).abs().mean()).backward()
        kd.bucket.pack_grads()
    zs = fresh_latents(B)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            slice_step(zs)
    torch.cuda.current_stream(dev).wait_stream(side)
    cx.barrier_sync()
    if use_graph:
        sg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(sg):
            slice_step(zs)
        run_slice = sg.replay
    else:
        run_slice = lambda: slice_step(zs)
    run_slice()
    cx.barrier_sync()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(args.steps):
        run_slice()
    s1.record()
    cx.barrier_sync()
    slice_ms = s0.elapsed_time(s1) / args.steps

    # ---------------- the other scaling mode, same run (N > 1 only: at N = 1 the two coincide)
    other = None
    if world > 1 and not args.no_second_mode:
        B2 = GLOBAL_BATCH if strong else GLOBAL_BATCH // world
        del kd.graph
        kd.graph = None
        ms2, _, _, _, _ = measure(kd, B2, args.steps, args.warmup, False)
        other = {'scaling': 'weak' if strong else 'strong', 'per_gpu_batch': B2, 'global_batch': B2 * world,
                 'value': world * B2 * args.steps / (ms2 / 1e3), 'unit': 'images/s', 'ms_per_step': ms2 / args.steps}

    # ---------------- the rounds-1/2 workload ("KD-like": no LPIPS, no parser, constant mask), same modules, same run
    kd_like = None
    if full and not args.no_kd_like:
        kd.percept_loss, kd.parsing_net, kd.mask = None, None, synthetic_mask(size, dev)
        del kd.graph
        kd.graph = None
        ms3, _, _, _, _ = measure(kd, B, args.steps, args.warmup, False)
        kd_like = {'value': world * B * args.steps / (ms3 / 1e3), 'unit': 'images/s', 'ms_per_step': ms3 / args.steps,
                   'workload': kd_workload(size, 'kdlike')}
        kd.percept_loss, kd.parsing_net, kd.mask = percept, parser, None

    if rank != 0:
        return
    peaks = load_peaks()
    summ = prof.summary()
    kernels, layers = {}, {}
    for name, r in summ.items():
        if '@' in name:     # per-layer records: kept apart from the per-kernel aggregates
            layers[name] = {'launches_per_step': r['launches'] / prof_steps,
                            'avg_launch_us': r['ms'] / max(r['launches'], 1) * 1e3,
                            'tflops': (r['flops'] / (r['ms'] / 1e3) / 1e12) if r['flops'] and r['ms'] else None,
                            'gbs': (r['bytes'] / (r['ms'] / 1e3) / 1e9) if r['bytes'] and r['ms'] else None}
    summ = {k: v for k, v in summ.items() if '@' not in k}
    for name, r in summ.items():
        per = r['ms'] / max(r['launches'], 1)
        kernels[name] = {'launches_per_step': r['launches'] / prof_steps, 'ms_per_step': r['ms'] / prof_steps,
                         'avg_launch_us': per * 1e3,
                         'tflops': (r['flops'] / (r['ms'] / 1e3) / 1e12) if r['flops'] and r['ms'] else None,
                         'gbs': (r['bytes'] / (r['ms'] / 1e3) / 1e9) if r['bytes'] and r['ms'] else None}
    conv_names = [n for n in summ if n.startswith('conv_') or n.startswith('dconv') or n.startswith('lpips_conv')]
    dom = max(conv_names, key=lambda n: summ[n]['ms']) if conv_names else None
    roofline = None
    if dom:
        r = summ[dom]
        ach = r['flops'] / (r['ms'] / 1e3) / 1e12
        tf32 = 'algo1' in dom
        # The per-kernel events come from a short eager burst (a few steps, each kernel bracketed alone): the BURST
        # tensor peak applies.  TF32 peak = half the measured bf16 peak (same pipe, half the rate); fp32 SIMT kernels
        # are reported against the same number so the fraction shows the distance to the target.
        peak = peaks['bf16_tflops'] / 2
        traffic, traffic_src = None, None
        for tname in ('r2_traffic.json', 'r1_traffic.json'):
            tpath = os.path.join(ROOT, 'profiles', tname)
            if tf32 and os.path.exists(tpath):
                tj = json.load(open(tpath)).get('conv_tc_family')
                if tj:
                    traffic, traffic_src = tj['dram_bytes_per_launch'], tj['source']
                    break
        roofline = {'kernel': dom, 'bound': 'tensor', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s',
                    'frac': ach / peak, 'traffic': traffic, 'traffic_source': traffic_src,
                    'algorithmic_bytes_per_launch': r['bytes'] / max(r['launches'], 1),
                    'algorithmic_flops_per_launch': r['flops'] / max(r['launches'], 1),
                    'avg_launch_us': r['ms'] / max(r['launches'], 1) * 1e3,
                    'peak_source': f"{peaks['source']} bf16 BURST {peaks['bf16_tflops']} TF/s / 2 (TF32); kernels "
                                   f"timed alone in a {prof_steps}-step eager burst",
                    'precision': 'tf32 tcgen05' if tf32 else 'fp32 SIMT (no tensor pipe)'}
    hbm = {}
    for n in ('fir_nhwc',):
        if n in summ and summ[n]['ms']:
            g = summ[n]['bytes'] / (summ[n]['ms'] / 1e3) / 1e9
            hbm[n] = {'bound': 'hbm', 'achieved': g, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                      'frac': g / peaks['hbm_gbs'],
                      'note': 'all generator FIR launches of a step, 9x9 .. 257x257 (the small ones are launch-latency bound)'}
    # the largest FIR call of the step (Blur after the last up-conv) and the heaviest conv layers, each alone
    fl = {k: v for k, v in layers.items() if k.startswith('fir_nhwc@') and v['gbs']}
    if fl:
        k = max(fl, key=lambda k: fl[k]['avg_launch_us'])
        hbm['fir_nhwc[largest]'] = {'bound': 'hbm', 'layer': k.split('@')[1], 'achieved': fl[k]['gbs'],
                                    'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': fl[k]['gbs'] / peaks['hbm_gbs'],
                                    'avg_launch_us': fl[k]['avg_launch_us']}
    cl = {k: v for k, v in layers.items() if k.startswith(('conv_', 'dconv', 'lpips_conv')) and v['tflops']}
    top_layers = {k: cl[k] for k in sorted(cl, key=lambda k: -cl[k]['avg_launch_us'] * cl[k]['launches_per_step'])[:6]}

    cpu, ref_gpu = None, None
    if world == 1 and not args.no_cpu_baseline:
        ref = run_ref_step('cpu', size, 2, 2, 1, timeout=600, loss=args.loss)
        cpu = cpu_baseline_from(ref, 2, 2, 1) or {'unavailable': ref.get('unavailable')}
    if world == 1 and not args.no_ref_gpu:
        # the real competitor: the reference CUDA path -- its SIMT op/*.cu compiled for sm_100a +
        # cuDNN grouped convolutions (TF32 allowed, torch's default) -- on this same B200, full batch, CUDA events
        torch.cuda.empty_cache()
        rg = run_ref_step('cuda', size, GLOBAL_BATCH, 30 if size == 256 else 6, 8 if size == 256 else 2, timeout=900,
                          loss=args.loss)
        if 'unavailable' in rg:
            ref_gpu = rg
        else:
            ref_gpu = {'value': rg['images_per_s'], 'unit': 'images/s', 'ms_per_step': rg['sec_per_step'] * 1e3,
                       'batch': rg['batch'], 'steps': rg['steps'], 'warmup': rg['warmup'], 'tf32': rg['tf32'],
                       'speedup_of_this_repo': value / rg['images_per_s'],
                       'what': "the reference's own model.py + op/*.cu (sm_100a JIT build) + cuDNN on this B200, "
                               + ('its own KD_loss / lpips / BiSeNet / mask glue, same KD step'
                                  if rg.get('loss_kind') == 'full' else 'same KD-like step')
                               + ', eager (no CUDA graph: the reference has none and its mask glue syncs with the host)'}

    gf = GFLOP_PER_IMG[size]
    line = {
        'metric': f'images/sec KD step {size}px pruned StyleGAN2', 'value': value, 'unit': 'images/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_total / args.steps,
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'tf32' if cx.tf32 else 'f32', 'data': 'synthetic',
        'config': kd_config(size, world, args.loss) if strong
        else dict(kd_config(size, world, args.loss), global_batch=B * world),
        'config_detail': {'per_gpu_batch': B, 'conv_algo': 'tcgen05-tf32' if cx.tf32 else 'simt-fp32',
                   'launch': ('one CUDA graph per step' if use_graph else 'eager launches') + (', teacher forward on a parallel graph branch' if kd.teacher_stream is not None else ''),
                   'kernel_timing': f'CUDA events around each native launch over {prof_steps} eager steps in this run'},
        'e2e': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': 2 * B * 512 * 4 * world,
                'd2h_bytes_per_step': 4 * world},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': roofline,
        'roofline_hbm': hbm,
        'cpu_baseline': cpu,
        'reference_gpu': ref_gpu,
        'sustained': sustained,
        ('weak_scaling' if strong else 'strong_scaling'): other,
        'kd_like': kd_like,
        'generator_slice': {'ms_per_step': slice_ms, 'images_per_s': B / (slice_ms / 1e3),
                            'tflops': B * (3 * gf['student'] + gf['teacher']) / slice_ms,
                            'note': 'student f+b + teacher f only, rank 0'},
        'kernels': kernels,
        'top_conv_layers': top_layers,
    }
    if args.all_layers:
        line['all_layers'] = layers
    print(json.dumps(line))


def bench_fid(args, cx):
    """BASELINE configs[4]: get_fid.py's sampling loop (Evaluation/fid.py:19-38) -- 256px pruned generator, batch 64
    per step sharded over the ranks, truncation 1, fresh latents and noise every batch, under no_grad.  Generator
    only (Inception and the FID arithmetic are outside the path).  One CUDA graph per batch; e2e = pinned-host
    latents in, images back to pinned host memory."""
    torch = cx.torch
    import model
    dev, world, rank = cx.dev, cx.world, cx.rank
    size, gb = 256, 64
    assert gb % world == 0
    B = gb // world
    args.warmup = max(args.warmup, 3)
    gen = model.Generator(size, 512, 8, generator_net_shape=STUDENT_SHAPES[size]).to(dev).eval()
    cx.broadcast_module(gen)
    for p in gen.parameters():
        p.requires_grad_(False)
    z_static = torch.zeros(B, 512, device=dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(3):
            gen([z_static], truncation=1)
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    n0 = cx.lib.launch_count()
    with torch.cuda.graph(graph), torch.no_grad():
        img_static = gen([z_static], truncation=1)
    per_replay = cx.lib.launch_count() - n0

    def step(i):
        z_static.normal_()
        graph.replay()
    for _ in range(args.warmup):
        step(0)
    sampler = ClockSampler(cx.local) if rank == 0 else None
    cx.barrier_sync()
    if sampler:
        sampler.start()
    ms = cx.time_steps(step, args.steps)
    clocks = sampler.stop() if sampler else None
    value = gb * args.steps / (ms / 1e3)
    hz = [torch.randn(B, 512).pin_memory() for _ in range(8)]
    himg = torch.empty(B, 3, size, size).pin_memory()

    def e2e_step(i):
        z_static.copy_(hz[i % 8], non_blocking=True)
        graph.replay()
        himg.copy_(img_static, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    e2e_step(0)
    e2e_ms = cx.time_steps(e2e_step, args.steps)
    if rank != 0:
        return
    gf = GFLOP_PER_IMG[size]['student']
    print(json.dumps({
        'metric': 'images/sec get_fid.py sampling 256px pruned StyleGAN2', 'value': value, 'unit': 'images/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'tf32' if cx.tf32 else 'f32',
        'data': 'synthetic',
        'config': {'workload': 'get_fid.py sampling loop: 256px 70%-pruned generator forward, batch 64 per step '
                               '(sharded over ranks), truncation 1, fresh latents + noise per batch',
                   'global_batch': gb, 'per_gpu_batch': B, 'parallelism': f'dp{world}', 'launch': 'one CUDA graph per batch',
                   'l2': 'activations of a batch (>1 GB) exceed the 126 MB L2'},
        'e2e': {'value': gb * args.steps / (e2e_ms / 1e3), 'unit': 'images/s', 'h2d_bytes_per_step': gb * 512 * 4,
                'd2h_bytes_per_step': gb * 3 * size * size * 4},
        'gpu_launches': int(per_replay * args.steps), 'clocks': clocks,
        'generator_tflops': gb * gf / (ms / args.steps), 'roofline': None, 'cpu_baseline': None}))


def bench_saliency(args, cx):
    """BASELINE configs[2]: content-aware saliency, full 256px generator, batches of 8 latents (forward with the
    autograd graph, sparse +-1 cotangent, backward, per-channel scores).  One step = one batch of 8.  Reports both
    engines: exact-fp32 SIMT and the fp32-accurate tensor-pipe split (3xTF32) when available."""
    torch, config = cx.torch, cx.config
    import model
    from b200gan import saliency as S
    dev, rank, world = cx.dev, cx.rank, cx.world
    gen = model.Generator(256, 512, 8).to(dev)
    cx.broadcast_module(gen)
    bs = 8
    res = {}
    modes = [('simt_fp32', config.ALGO_SIMT_FP32)]
    if hasattr(config, 'ALGO_TCGEN05_3XTF32'):
        modes.append(('tcgen05_3xtf32', config.ALGO_TCGEN05_3XTF32))
    for name, algo in modes:
        def one(i):
            S.content_aware_scores(gen, bs, bs, 0.05, dev, seed=100 + i + 1000 * rank, algo=algo) \
                if 'algo' in S.content_aware_scores.__code__.co_varnames else \
                S.content_aware_scores(gen, bs, bs, 0.05, dev, seed=100 + i + 1000 * rank)
        for i in range(max(1, min(args.warmup, 2))):
            one(i)
        ms = cx.time_steps(one, args.steps)
        res[name] = {'ms_per_batch_of_8': ms / args.steps, 'latents_per_s': world * bs * args.steps / (ms / 1e3),
                     'tflops': world * bs * 3 * GFLOP_PER_IMG[256]['teacher'] / (ms / args.steps)}
    if rank != 0:
        return
    best = max(res, key=lambda k: res[k]['latents_per_s'])
    print(json.dumps({
        'metric': 'latents/sec content-aware saliency 256px full StyleGAN2', 'value': res[best]['latents_per_s'],
        'unit': 'latents/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': res[best]['ms_per_batch_of_8'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'saliency pass (Util/content_aware_pruning.py:200-249): full 256px generator fwd+bwd '
                               'per batch of 8 latents, whole batches per rank, host-side mask/noise included',
                   'engine': best, 'parallelism': f'dp{world}'},
        'engines': res, 'roofline': None, 'cpu_baseline': None, 'gpu_launches': None, 'e2e': None}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='kd256', choices=['kd256', 'kd1024', 'fid256', 'saliency256'])
    ap.add_argument('--size', type=int, default=None, choices=[256, 1024])
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'],
                    help='strong: global batch 16 sharded over the ranks (north star); weak: 16 per GPU')
    ap.add_argument('--algo', default='auto', choices=['auto', 'simt', 'tc'])
    ap.add_argument('--sustained-steps', type=int, default=200)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-ref-gpu', action='store_true')
    ap.add_argument('--no-second-mode', action='store_true')
    ap.add_argument('--loss', default='full', choices=['full', 'kdlike'],
                    help='full: train.py:280-308 with parser + LPIPS (configs[1]); kdlike: the rounds-1/2 workload')
    ap.add_argument('--no-kd-like', action='store_true', help='skip the secondary measurement of the KD-like workload')
    ap.add_argument('--all-layers', action='store_true', help='add every per-layer kernel record to the JSON line')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of one CUDA graph per step')
    args = ap.parse_args()
    if args.size is None:
        args.size = 1024 if args.config == 'kd1024' else 256
    if args.impl == 'reference':
        return run_reference(args)
    cx = Ctx(args)
    if args.config in ('kd256', 'kd1024'):
        bench_kd(args, cx)
    elif args.config == 'fid256':
        bench_fid(args, cx)
    else:
        bench_saliency(args, cx)
    import torch
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
