#!/usr/bin/env python
"""Benchmark of the KD retraining generator step (BASELINE.json metric: images/sec, 256px 70%-pruned
StyleGAN2 student, full-size teacher, batch 16 per GPU) -- see DESIGN.md "Measurement".

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on the host cores

One step = train.py:280-308 restricted to the hot path and its direct consumers: student forward
(RGB list) + discriminator forward + GAN loss, teacher forward, masked L1 KD loss, backward, gradient
all-reduce over ranks, fused Adam.  LPIPS-VGG / BiSeNet are third-party networks outside the path
(their weights are not available offline) and are excluded on both arms.
Prints ONE JSON line on rank 0.
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
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
    del os.environ['NCCL_DEBUG']           # both print NCCL's version banner on stdout: rank 0 prints ONE JSON line

import torch  # noqa: E402

STUDENT_SHAPES = {256: [154] * 10 + [77, 77, 39, 39],
                  1024: [154] * 10 + [77, 77, 39, 39, 20, 20, 10, 10]}
# conv FLOPs per image, reference convention (Util/Calculators.py:16-61 x2), SURVEY.md §8d
GFLOP_PER_IMG = {256: {'student': 8.240, 'teacher': 90.236}, 1024: {'student': 13.973, 'teacher': 148.520}}


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


def synthetic_mask(size, device):
    """Stand-in for the BiSeNet face mask (Util/content_aware_pruning.py:61-117): centred ellipse."""
    yy, xx = torch.meshgrid(torch.arange(size, device=device), torch.arange(size, device=device), indexing='ij')
    c = (size - 1) / 2
    return ((((yy - c) / (0.42 * size)) ** 2 + ((xx - c) / (0.34 * size)) ** 2) <= 1).float().view(1, 1, size, size)


def cpu_reference_step_time(size, batch, steps, warmup, seed=0):
    """The reference algorithm's CPU path (oracle port, fp32, all host threads) on a bounded sample of
    the same workload: `batch` images of the same KD-like step.  Returns (images/s, seconds/step, cores)."""
    import model  # only to draw identical synthetic weights; no CUDA call is made here
    from oracle import stylegan2_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(seed)
    teacher = model.Generator(size, 512, 8)
    student = model.Generator(size, 512, 8, generator_net_shape=STUDENT_SHAPES[size])
    disc = model.Discriminator(size)
    tp = {k: v.detach() for k, v in teacher.state_dict().items()}
    sp = {k: v.detach() for k, v in student.state_dict().items()}
    dp = {k: v.detach() for k, v in disc.state_dict().items()}
    mask = synthetic_mask(size, 'cpu')
    times = []
    for it in range(warmup + steps):
        z = [torch.randn(batch, 512), torch.randn(batch, 512)]
        s_noise = [torch.randn(batch, 1, n.shape[2], n.shape[3]) for n in student.make_noise()]
        t_noise = [torch.randn(batch, 1, n.shape[2], n.shape[3]) for n in teacher.make_noise()]
        t0 = time.perf_counter()
        O.kd_step(sp, tp, dp, size, z, s_noise, t_noise, 5, mask)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return batch / sec, sec, cores


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample_b = 2
    ips, sec, cores = cpu_reference_step_time(args.size, sample_b, args.steps, max(1, min(args.warmup, 1)))
    line = {
        'impl': 'reference', 'metric': f'images/sec KD step {args.size}px pruned StyleGAN2', 'value': ips,
        'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'KD-like generator step {args.size}px, 70%-pruned student + full teacher + D, '
                               f'bounded sample: batch {sample_b} per step (the GPU arm runs batch {args.batch}/GPU)'},
        'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                         'sample': f'oracle.kd_step, batch {sample_b}, {args.steps} timed steps, torch CPU fp32, '
                                   f'{cores} threads'},
        'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--size', type=int, default=256, choices=[256, 1024])
    ap.add_argument('--batch', type=int, default=16, help='images per GPU per step (weak scaling)')
    ap.add_argument('--algo', default='auto', choices=['auto', 'simt', 'tc'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of one CUDA graph per step')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    from b200gan import dist as D, config, _lib
    from b200gan.kd import KDStep
    import model

    local = D.init_from_env()
    rank, world = D.get_rank(), D.get_world_size()
    assert torch.cuda.is_available(), 'bench.py (impl b200) needs a CUDA device; there is no CPU fallback'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)

    algo = {'simt': config.ALGO_SIMT_FP32, 'tc': config.ALGO_TCGEN05_TF32}.get(args.algo)
    if algo is None:
        algo = config.best_available_algo()
    config.set_default_algo(algo)

    torch.manual_seed(1234 + rank)
    size, B = args.size, args.batch
    teacher = model.Generator(size, 512, 8).to(dev)
    student = model.Generator(size, 512, 8, generator_net_shape=STUDENT_SHAPES[size]).to(dev)
    disc = model.Discriminator(size).to(dev)
    if world > 1:   # identical replicas on every rank
        for m in (teacher, student, disc):
            for t in list(m.parameters()) + list(m.buffers()):
                torch.distributed.broadcast(t.data, 0)
    kd = KDStep(student, teacher, disc, mask=synthetic_mask(size, dev))
    inject = 5

    def fresh_latents():
        return [torch.randn(B, 512, device=dev), torch.randn(B, 512, device=dev)]

    def barrier_sync():
        D.synchronize()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident timing (value)
    use_graph = not args.no_graph
    if use_graph:
        kd.capture(B, inject)            # includes 3 eager warm-up steps on a side stream
        run_step = lambda z: kd.step_graphed(z)
    else:
        run_step = lambda z: kd.step(z, inject)
    for _ in range(args.warmup):
        run_step(fresh_latents())
    lat = [fresh_latents() for _ in range(args.steps)]
    sampler = ClockSampler(local)
    barrier_sync()
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        run_step(lat[i])
    e1.record()
    barrier_sync()
    launches = (kd.launches_per_replay * args.steps) if use_graph else (_lib.launch_count() - n0)
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * B * args.steps / (ms_total / 1e3)

    # ---------------- end-to-end through the public API: pinned host latents in, loss out, every step
    hz = [[torch.randn(B, 512).pin_memory(), torch.randn(B, 512).pin_memory()] for _ in range(args.steps)]
    for _ in range(2):
        kd.step_from_host(hz[0], inject)
    barrier_sync()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        kd.step_from_host(hz[i], inject)
    f1.record()
    barrier_sync()
    ems = torch.tensor([f0.elapsed_time(f1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ems, op=torch.distributed.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(ems.item()) / 1e3)

    # ---------------- per-kernel CUDA events (roofline): eager steps, events on the launching stream
    prof = config.KernelProfiler()
    overlap_stream, kd.teacher_stream = kd.teacher_stream, None   # one stream: a kernel's events bracket only itself
    kd.step(lat[0], inject)
    barrier_sync()
    config.set_profiler(prof)
    prof_steps = min(args.steps, 5)
    for i in range(prof_steps):
        kd.step(lat[i], inject)
    barrier_sync()
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
    zs = fresh_latents()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            slice_step(zs)
    torch.cuda.current_stream(dev).wait_stream(side)
    barrier_sync()
    if use_graph:
        sg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(sg):
            slice_step(zs)
        run_slice = sg.replay
    else:
        run_slice = lambda: slice_step(zs)
    run_slice()
    barrier_sync()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(args.steps):
        run_slice()
    s1.record()
    barrier_sync()
    slice_ms = s0.elapsed_time(s1) / args.steps

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
    conv_names = [n for n in summ if n.startswith('conv_')]
    dom = max(conv_names, key=lambda n: summ[n]['ms']) if conv_names else None
    roofline = None
    if dom:
        r = summ[dom]
        ach = r['flops'] / (r['ms'] / 1e3) / 1e12
        tf32 = 'algo1' in dom
        # TF32 tensor peak = half the measured bf16 peak (same pipe, half the rate); fp32 SIMT kernels are
        # reported against the same tensor-pipe number so the fraction shows the distance to the target
        peak = peaks['bf16_tflops_sustained'] / 2
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
        if tf32 and os.path.exists(tpath):
            tj = json.load(open(tpath)).get('conv_tc_family')
            if tj:
                traffic, traffic_src = tj['dram_bytes_per_launch'], tj['source']
        roofline = {'kernel': dom, 'bound': 'tensor', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s',
                    'frac': ach / peak, 'traffic': traffic, 'traffic_source': traffic_src,
                    'algorithmic_bytes_per_launch': r['bytes'] / max(r['launches'], 1),
                    'algorithmic_flops_per_launch': r['flops'] / max(r['launches'], 1),
                    'avg_launch_us': r['ms'] / max(r['launches'], 1) * 1e3,
                    'peak_source': f"{peaks['source']} bf16 sustained {peaks['bf16_tflops_sustained']} TF/s / 2 (TF32)",
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
    cl = {k: v for k, v in layers.items() if k.startswith('conv_') and v['tflops']}
    top_layers = {k: cl[k] for k in sorted(cl, key=lambda k: -cl[k]['avg_launch_us'] * cl[k]['launches_per_step'])[:6]}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ips, sec, cores = cpu_reference_step_time(size, 2, 2, 1)
        cpu = {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
               'sample': f'oracle.kd_step (CPU restatement of the reference path), batch 2, 1 warm-up + 2 timed steps, '
                         f'torch CPU fp32, {cores} threads, {sec:.2f} s/step'}

    gf = GFLOP_PER_IMG[size]
    line = {
        'metric': f'images/sec KD step {size}px pruned StyleGAN2', 'value': value, 'unit': 'images/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'tf32' if algo == config.ALGO_TCGEN05_TF32 else 'f32', 'data': 'synthetic',
        'config': {'workload': f'KD-like generator step {size}px: 70%-pruned student {STUDENT_SHAPES[size][0]}/'
                               f'{STUDENT_SHAPES[size][-1]}ch f+b, full teacher f, discriminator f+dgrad, masked L1 KD, '
                               f'grad all-reduce, fused Adam; batch {B}/GPU; style mixing inject_index={inject}',
                   'global_batch': B * world, 'parallelism': f'dp{world}',
                   'l2': 'per-step working set (>=4 GB of activations) exceeds the 126 MB L2; no explicit flush',
                   'conv_algo': 'tcgen05-tf32' if algo == config.ALGO_TCGEN05_TF32 else 'simt-fp32',
                   'launch': ('one CUDA graph per step' if use_graph else 'eager launches') + (', teacher forward on a parallel graph branch' if kd.teacher_stream is not None else ''),
                   'kernel_timing': f'CUDA events around each native launch over {prof_steps} eager steps in this run'},
        'e2e': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': 2 * B * 512 * 4,
                'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': roofline,
        'roofline_hbm': hbm,
        'cpu_baseline': cpu,
        'generator_slice': {'ms_per_step': slice_ms, 'images_per_s': B / (slice_ms / 1e3),
                            'tflops': B * (3 * gf['student'] + gf['teacher']) / slice_ms,
                            'note': 'student f+b + teacher f only, rank 0'},
        'kernels': kernels,
        'top_conv_layers': top_layers,
    }
    print(json.dumps(line))


if __name__ == '__main__':
    main()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
