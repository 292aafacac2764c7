"""profiles/r2_traffic.json + profiles/r2_summary.md from the artefacts brought back from the GPU box.

Inputs under profiles/ (copied from gpurun_out/ by hand, named per round):
  r2_bench_1gpu.json             final `python bench.py` line, complete KD step (parser + LPIPS), 1 GPU
  r2_bench_reference_arm.json    `python bench.py --impl reference` line (the reference's own CPU code)
  r2_scale_full.json             complete step at N = 2 / 8 (strong + weak + KD-like side figure)
  r2_scale.json                  KD-like workload at N = 1 / 2 / 4 / 8 (mid-round build)
  r2_side_configs.json           BASELINE configs[2], [3], [4]
  r2_launches_kdstep_b16.csv     ncu launch list (time + DRAM bytes) of one eager complete step, batch 16
  r2_launches_kdlike_b16.csv / _b2.csv   the same for the KD-like step (mid-round build), batch 16 / per-rank batch 2
  r2_ncu_full_layers.csv         ncu --set full on isolated layer shapes (compact)
"""
import collections
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, 'profiles')


def launches(path):
    rows = list(csv.DictReader([l for l in open(path) if not l.startswith('==')]))
    per = collections.defaultdict(dict)
    for r in rows:
        per[r['ID']]['name'] = r['Kernel Name'].split('(')[0]
        v = float(r['Metric Value'].replace(',', ''))
        u, m = r['Metric Unit'], r['Metric Name']
        if m == 'gpu__time_duration.sum':
            v = v / 1e3 if u in ('nsecond', 'ns') else v * 1e3 if u in ('msecond', 'ms') else v
        else:
            v = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1) * v
        per[r['ID']][m] = v
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in per.values():
        a = agg[d['name']]
        a[0] += 1
        a[1] += d.get('gpu__time_duration.sum', 0)
        a[2] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    return len(per), agg


def table(agg, top=24):
    tot = sum(a[1] for a in agg.values())
    out = ['| kernel | launches | us | share | DRAM GB | GB/s |', '|---|---|---|---|---|---|']
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f'| `{k[:70]}` | {a[0]} | {a[1]:.0f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e9:.3f} | {a[2] / max(a[1], 1e-9) / 1e3:.0f} |')
    return out


def load(name):
    path = os.path.join(P, name)
    return json.load(open(path)) if os.path.exists(path) else None


n16, a16 = launches(os.path.join(P, 'r2_launches_kdstep_b16.csv'))
conv = [k for k in a16 if 'conv_tc' in k]
cl = sum(a16[k][0] for k in conv)
cb = sum(a16[k][2] for k in conv)
ct = sum(a16[k][1] for k in conv)
tot16 = sum(a[1] for a in a16.values())
traffic = {'conv_tc_family': {'dram_bytes_per_launch': cb / cl, 'launches': cl, 'dram_bytes_total': cb,
                              'share_of_step_time': ct / tot16,
                              'source': 'profiles/r2_launches_kdstep_b16.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, one '
                                        'complete KD step, conv_tc_persist + conv_tc_halo + conv_tc kernels = generator, discriminator and '
                                        'VGG16 convolutions)'},
           'per_kernel': {k: {'launches': a[0], 'us': a[1], 'dram_bytes': a[2]} for k, a in a16.items() if a[1] > 50}}
json.dump(traffic, open(os.path.join(P, 'r2_traffic.json'), 'w'), indent=1)

b = load('r2_bench_1gpu.json')
r = b['roofline']
out = ['# r2 profile summary (B200, sm_100a)', '',
       '## bench.py, 1 GPU (driver-style run `python bench.py`, not under a profiler): the COMPLETE KD step', '',
       f"workload: {b['config']['workload']}", '',
       f"* value **{b['value']:.1f} images/s** ({b['ms_per_step']:.2f} ms/step, global batch 16), e2e {b['e2e']['value']:.1f}; sustained over "
       f"{b['sustained']['steps']} steps {b['sustained']['value']:.1f} images/s (clocks {b['sustained']['clocks']}); {b['gpu_launches'] // b['steps']} native launches per step; "
       f"generator slice {b['generator_slice']['ms_per_step']:.2f} ms ({b['generator_slice']['tflops']:.0f} TFLOP/s by the reference FLOP convention)"]
if b.get('kd_like'):
    out.append(f"* the rounds-1/2 workload in the same run (no LPIPS, no parser): {b['kd_like']['value']:.1f} images/s ({b['kd_like']['ms_per_step']:.2f} ms/step)")
if b.get('reference_gpu') and 'value' in b['reference_gpu']:
    g = b['reference_gpu']
    out.append(f"* reference CUDA path on the same B200 ({g['what']}): **{g['value']:.1f} images/s** ({g['ms_per_step']:.1f} ms/step) -> this repo {g['speedup_of_this_repo']:.2f}x")
if b.get('cpu_baseline') and 'value' in b['cpu_baseline']:
    out.append(f"* reference CPU path ({b['cpu_baseline']['cores']} host threads, batch 2, its own KD_loss / lpips / BiSeNet): {b['cpu_baseline']['value']:.2f} images/s")
out.append(f"* roofline (dominant conv op by time, CUDA events): {r['kernel']} {r['achieved']:.0f} TFLOP/s = {100 * r['frac']:.1f}% of {r['peak']:.0f} ({r['peak_source']})")
for k, v in b['roofline_hbm'].items():
    out.append(f"* {k}: {v['achieved']:.0f} GB/s = {100 * v['frac']:.1f}% of measured HBM {v['peak']}")
out += ['', '| op (events) | launches/step | ms/step | TFLOP/s | GB/s |', '|---|---|---|---|---|']
for k, v in sorted(b['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step']):
    out.append(f"| {k} | {v['launches_per_step']:.0f} | {v['ms_per_step']:.3f} | {(v['tflops'] or 0):.0f} | {(v['gbs'] or 0):.0f} |")
out += ['', '| heaviest layers | launches | us | TFLOP/s |', '|---|---|---|---|']
for k, v in b['top_conv_layers'].items():
    out.append(f"| {k} | {v['launches_per_step']:.0f} | {v['avg_launch_us']:.0f} | {(v['tflops'] or 0):.0f} |")

sf = load('r2_scale_full.json')
if sf:
    out += ['', '## scaling of the complete step, one box (torchrun, NCCL; `profiles/r2_scale_full.json`)', '',
            '| N | strong: images/s (global batch 16) | ms/step | efficiency vs N=1 | weak: images/s (16 per GPU) | KD-like strong images/s |', '|---|---|---|---|---|---|']
    v1 = b['value']
    out.append(f"| 1 | {v1:.0f} | {b['ms_per_step']:.2f} | 1.00 | {v1:.0f} | {b['kd_like']['value'] if b.get('kd_like') else 0:.0f} |")
    for n in sorted(sf, key=int):
        d = sf[n]
        w = d.get('weak_scaling') or {}
        out.append(f"| {n} | {d['value']:.0f} | {d['ms_per_step']:.2f} | {d['value'] / v1 / int(n):.2f} | {w.get('value', 0):.0f} | "
                   f"{(d.get('kd_like') or {}).get('value', 0):.0f} |")
sc = load('r2_scale.json')
if sc:
    sc = sc['kd256_strong_and_weak']
    out += ['', '## scaling of the KD-like workload (mid-round build; `profiles/r2_scale.json`)', '',
            '| N | strong: images/s (global batch 16) | ms/step | efficiency | weak: images/s (16 per GPU) | efficiency |', '|---|---|---|---|---|---|']
    v1 = sc['1']['value']
    for n in ('1', '2', '4', '8'):
        d = sc[n]
        w = d.get('weak_scaling')
        out.append(f"| {n} | {d['value']:.0f} | {d['ms_per_step']:.2f} | {d['value'] / v1 / int(n):.2f} | "
                   f"{(w['value'] if w else d['value']):.0f} | {((w['value'] if w else d['value']) / v1 / int(n)):.2f} |")
side = load('r2_side_configs.json')
if side:
    out += ['', '## other BASELINE configs (`profiles/r2_side_configs.json`)', '']
    for k, v in side.items():
        out.append(f"* {k}: {v['value']:.1f} {v['unit']} ({v.get('ms_per_step') and round(v['ms_per_step'], 2)} ms/step)" +
                   (f", weak {v['weak_scaling']['value']:.0f}" if v.get('weak_scaling') else '') +
                   (f" -- {v['note']}" if v.get('note') else ''))
out += ['', f'## ncu launch list of one eager COMPLETE KD step, batch 16 ({n16} launches, {tot16 / 1e3:.2f} ms serialized; `profiles/r2_launches_kdstep_b16.csv`)',
        '', '(cold-cache, serialized, single stream: compare shares, not absolutes; `sm80_xmma*` / `cutlass3x*` / `at::*` rows are the face',
        "parser's library convolutions and glue, everything else is this repo's kernels)", '']
out += table(a16, 40)
for tag, title in (('r2_launches_kdlike_b16.csv', 'KD-like step (no LPIPS / parser), batch 16, mid-round build'),
                   ('r2_launches_kdlike_b2.csv', 'KD-like step at batch 2 per GPU (the per-rank work of 8-GPU strong scaling), mid-round build')):
    if os.path.exists(os.path.join(P, tag)):
        n, a = launches(os.path.join(P, tag))
        out += ['', f'## {title} ({n} launches, {sum(x[1] for x in a.values()) / 1e3:.2f} ms serialized; `profiles/{tag}`)', '']
        out += table(a, 14)
out += ['', '## `ncu --set full` on isolated layer shapes (`profiles/r2_ncu_full_layers.csv`, scripts/profile_layers.py, batch 16)', '',
        '```'] + open(os.path.join(P, 'r2_ncu_full_layers.csv')).read().strip().splitlines() + ['```',
        '', 'Rows in launch order: teacher 512ch@64^2 (persistent kernel: tensor pipe 85.7 % active, r1: 62.7 %), teacher 128ch@256^2 (row-mode halo kernel: 58.4 %,',
        'r1: 42.6 % at 720 us), up-conv 256->128 (multi-phase persistent: 51 %) and its Blur (FIR row ring, DRAM 55.7 % of the ncu peak), then the student',
        'layers 154@64^2 (fwd / dgrad / wgrad), 39@256^2 and the 77->39 up-conv with their gradients.  SASS evidence: `profiles/r2_sass_summary.txt`.',
        '', 'The LPIPS / mask-glue bandwidth kernels have no `--set full` capture (the one attempt produced a report larger than the 64 MiB',
        'that travels back from the GPU box); their achieved DRAM rate is in the launch list above (time + dram bytes per launch):',
        'relu_pool_bwd 6.4 TB/s, maxpool2 6.3, relu_mask 5.9, lpips_head fwd/bwd 5.8-6.3, conv1_1 forward / backward 1.3 / 1.4 (LSU-bound).']
open(os.path.join(P, 'r2_summary.md'), 'w').write('\n'.join(out) + '\n')
print('\n'.join(out[:40]))
