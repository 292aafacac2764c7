"""profiles/r2_traffic.json + profiles/r2_summary.md from the artefacts brought back from the GPU box."""
import collections, csv, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, 'profiles')


def launches(path):
    rows = list(csv.DictReader([l for l in open(path) if not l.startswith('==')]))
    per = collections.defaultdict(dict)
    for r in rows:
        per[r['ID']]['name'] = r['Kernel Name'].split('(')[0]
        v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']; m = r['Metric Name']
        if m == 'gpu__time_duration.sum':
            v = v / 1e3 if u in ('nsecond', 'ns') else v * 1e3 if u in ('msecond', 'ms') else v
        else:
            v = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1) * v
        per[r['ID']][m] = v
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in per.values():
        a = agg[d['name']]
        a[0] += 1; a[1] += d.get('gpu__time_duration.sum', 0)
        a[2] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    return len(per), agg


def table(agg, top=24):
    tot = sum(a[1] for a in agg.values())
    out = ['| kernel | launches | us | share | DRAM GB | GB/s |', '|---|---|---|---|---|---|']
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f'| `{k[:70]}` | {a[0]} | {a[1]:.0f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e9:.3f} | {a[2] / max(a[1], 1e-9) / 1e3:.0f} |')
    return out


n16, a16 = launches(os.path.join(P, 'r2_launches_kdstep_b16.csv'))
n2, a2 = launches(os.path.join(P, 'r2_launches_kdstep_b2.csv'))
conv = [k for k in a16 if 'conv_tc' in k]
cl = sum(a16[k][0] for k in conv); cb = sum(a16[k][2] for k in conv); ct = sum(a16[k][1] for k in conv)
tot16 = sum(a[1] for a in a16.values())
traffic = {'conv_tc_family': {'dram_bytes_per_launch': cb / cl, 'launches': cl, 'dram_bytes_total': cb,
                              'share_of_step_time': ct / tot16,
                              'source': 'profiles/r2_launches_kdstep_b16.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, one KD step, '
                                        'conv_tc_persist + conv_tc_halo + conv_tc kernels = generator and discriminator convolutions)'},
           'per_kernel': {k: {'launches': a[0], 'us': a[1], 'dram_bytes': a[2]} for k, a in a16.items() if a[1] > 50}}
json.dump(traffic, open(os.path.join(P, 'r2_traffic.json'), 'w'), indent=1)

b = json.load(open(os.path.join(P, 'r2_bench_1gpu.json')))
sc = json.load(open(os.path.join(P, 'r2_scale.json')))['kd256_strong_and_weak']
side = json.load(open(os.path.join(P, 'r2_side_configs.json')))
r = b['roofline']
out = ['# r2 profile summary (B200, sm_100a)', '',
       '## bench.py, 1 GPU (driver-style run `python bench.py --steps 20 --warmup 5`, not under a profiler)', '',
       f"* value **{b['value']:.1f} images/s** ({b['ms_per_step']:.2f} ms/step, global batch 16), e2e {b['e2e']['value']:.1f}; sustained over "
       f"{b['sustained']['steps']} steps {b['sustained']['value']:.1f} images/s (clocks {b['sustained']['clocks']}); {b['gpu_launches'] // b['steps']} native launches per step; "
       f"generator slice {b['generator_slice']['ms_per_step']:.2f} ms ({b['generator_slice']['tflops']:.0f} TFLOP/s by the reference FLOP convention)",
       f"* reference CUDA path on the same B200 (its op/*.cu built for sm_100a + cuDNN, batch 16, eager): **{b['reference_gpu']['value']:.1f} images/s** "
       f"({b['reference_gpu']['ms_per_step']:.1f} ms/step) -> this repo {b['reference_gpu']['speedup_of_this_repo']:.2f}x",
       f"* reference CPU path (its own model.py + op fallbacks, {b['cpu_baseline']['cores']} host threads, batch 2): {b['cpu_baseline']['value']:.2f} images/s",
       f"* roofline (dominant conv op by time, CUDA events): {r['kernel']} {r['achieved']:.0f} TFLOP/s = {100 * r['frac']:.1f}% of {r['peak']:.0f} ({r['peak_source']})"]
for k, v in b['roofline_hbm'].items():
    out.append(f"* {k}: {v['achieved']:.0f} GB/s = {100 * v['frac']:.1f}% of measured HBM {v['peak']}")
out += ['', '| op (events) | launches/step | ms/step | TFLOP/s | GB/s |', '|---|---|---|---|---|']
for k, v in sorted(b['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step']):
    out.append(f"| {k} | {v['launches_per_step']:.0f} | {v['ms_per_step']:.3f} | {(v['tflops'] or 0):.0f} | {(v['gbs'] or 0):.0f} |")
out += ['', '| heaviest layers | launches | us | TFLOP/s |', '|---|---|---|---|']
for k, v in b['top_conv_layers'].items():
    out.append(f"| {k} | {v['launches_per_step']:.0f} | {v['avg_launch_us']:.0f} | {(v['tflops'] or 0):.0f} |")
out += ['', '## scaling, one 8-GPU box (torchrun, NCCL; `profiles/r2_scale.json`)', '',
        '| N | strong: images/s (global batch 16) | ms/step | efficiency | weak: images/s (16 per GPU) | efficiency |', '|---|---|---|---|---|---|']
v1 = sc['1']['value']
for n in ('1', '2', '4', '8'):
    d = sc[n]; w = d.get('weak_scaling')
    out.append(f"| {n} | {d['value']:.0f} | {d['ms_per_step']:.2f} | {d['value'] / v1 / int(n):.2f} | "
               f"{(w['value'] if w else d['value']):.0f} | {((w['value'] if w else d['value']) / v1 / int(n)):.2f} |")
out += ['', '## other BASELINE configs (`profiles/r2_side_configs.json`)', '']
for k, v in side.items():
    out.append(f"* {k}: {v['value']:.1f} {v['unit']} ({v.get('ms_per_step') and round(v['ms_per_step'], 2)} ms/step)" +
               (f", weak {v['weak_scaling']['value']:.0f}" if v.get('weak_scaling') else ''))
out += ['', f'## ncu launch list of one eager KD step, batch 16 ({n16} launches, {tot16 / 1e3:.2f} ms serialized; `profiles/r2_launches_kdstep_b16.csv`)', '']
out += table(a16)
out += ['', f'## same at batch 2 per GPU (the per-rank work of 8-GPU strong scaling; {n2} launches, {sum(a[1] for a in a2.values()) / 1e3:.2f} ms serialized)', '']
out += table(a2, 14)
out += ['', '## `ncu --set full` on isolated layer shapes (`profiles/r2_ncu_full_layers.csv`, scripts/profile_layers.py, batch 16)', '',
        '```'] + open(os.path.join(P, 'r2_ncu_full_layers.csv')).read().strip().splitlines() + ['```',
        '', 'Rows in launch order: teacher 512ch@64^2 (persistent kernel: tensor pipe 85.7 % active, r1: 62.7 %), teacher 128ch@256^2 (row-mode halo kernel: 58.4 %,',
        'r1: 42.6 % at 720 us), up-conv 256->128 (multi-phase persistent: 51 %) and its Blur (FIR row ring, DRAM 55.7 % of the ncu peak), then the student',
        'layers 154@64^2 (fwd / dgrad / wgrad), 39@256^2 and the 77->39 up-conv with their gradients.  SASS evidence: `profiles/r2_sass_summary.txt`.']
open(os.path.join(P, 'r2_summary.md'), 'w').write('\n'.join(out) + '\n')
print('\n'.join(out[:60]))
