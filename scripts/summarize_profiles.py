"""Turn the ncu CSVs brought back from the GPU box into profiles/<round>_summary.md."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, 'profiles')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'


def us(row):
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    return v / 1e3 if u in ('nsecond', 'ns') else v * 1e3 if u in ('msecond', 'ms') else v


out = [f'# {tag} profile summary (B200, sm_100a)\n']
bench = os.path.join(P, f'{tag}_bench_1gpu.json')
if os.path.exists(bench):
    d = json.load(open(bench))
    out.append('## bench.py (1 GPU, not under a profiler)\n')
    out.append(f"* value **{d['value']:.1f} {d['unit']}** ({d['ms_per_step']:.2f} ms/step), e2e {d['e2e']['value']:.1f}; "
               f"generator slice {d['generator_slice']['ms_per_step']:.2f} ms ({d['generator_slice']['tflops']:.1f} TFLOP/s); "
               f"cpu_baseline {d['cpu_baseline']['value']:.2f} img/s on {d['cpu_baseline']['cores']} cores; clocks {d['clocks']}")
    r = d['roofline']
    out.append(f"* roofline (dominant conv kernel, CUDA events): {r['kernel']} {r['achieved']:.1f} {r['unit']} = "
               f"{100 * r['frac']:.1f}% of {r['peak']:.1f} ({r['peak_source']})")
    for k, v in d.get('roofline_hbm', {}).items():
        out.append(f"* {k}: {v['achieved']:.0f} GB/s = {100 * v['frac']:.1f}% of measured HBM {v['peak']}")
    out.append('\n| kernel (events) | launches/step | ms/step | TFLOP/s | GB/s |\n|---|---|---|---|---|')
    for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step']):
        out.append(f"| {k} | {v['launches_per_step']:.0f} | {v['ms_per_step']:.3f} | "
                   f"{(v['tflops'] or 0):.1f} | {(v['gbs'] or 0):.0f} |")

lst = os.path.join(P, f'{tag}_launches_kdstep.csv')
if os.path.exists(lst):
    rows = csv.DictReader([l for l in open(lst) if not l.startswith('==')])
    agg = collections.defaultdict(lambda: [0, 0.0])
    dram = collections.defaultdict(float)
    tot = 0.0
    for row in rows:
        n = row['Kernel Name'].replace('void ', '')[:72]
        try:
            if row.get('Metric Name', 'gpu__time_duration.sum') == 'gpu__time_duration.sum':
                t = us(row)
                agg[n][0] += 1
                agg[n][1] += t
                tot += t
            elif row['Metric Name'].startswith('dram__bytes'):
                mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(row['Metric Unit'], 1)
                dram[n] += float(row['Metric Value'].replace(',', '')) * mult
        except (ValueError, KeyError):
            continue
    if dram:
        fam = [k for k in agg if 'conv_tc' in k]
        traffic = {'conv_tc_family': {'launches': sum(agg[k][0] for k in fam),
                                      'dram_bytes_per_launch': sum(dram[k] for k in fam) / max(1, sum(agg[k][0] for k in fam)),
                                      'source': f'{tag}_launches_kdstep.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum, one KD step)'}}
        for k in agg:
            traffic[k.split('(')[0]] = {'launches': agg[k][0], 'dram_bytes_per_launch': dram[k] / agg[k][0]}
        json.dump(traffic, open(os.path.join(P, f'{tag}_traffic.json'), 'w'), indent=1)
    out.append(f'\n## ncu launch list of one KD step (`--metrics gpu__time_duration.sum --clock-control none`, '
               f'cold-cache serialised: compare shares)\n\ntotal {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches\n')
    out.append('| kernel | launches | us | share | DRAM MB |\n|---|---|---|---|---|')
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        out.append(f'| `{k}` | {n} | {t:.0f} | {100 * t / tot:.1f}% | {dram.get(k, 0) / 1e6:.0f} |')

raw = os.path.join(P, f'{tag}_ncu_full_layers_raw.csv')
if os.path.exists(raw):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size']
    out.append('\n## `ncu --set full` on isolated KD-step layer shapes (scripts/profile_layers.py, batch 16)\n')
    out.append('| kernel | grid | time | dram read | dram write | dram % | tensor pipe % | warps active % | regs |\n|---|---|---|---|---|---|---|---|---|')
    for r in rows[2:]:
        name = r[idx['Kernel Name']].split('(')[0][-40:]
        if 'modulate' in name and float(r[idx['gpu__time_duration.sum']]) < 0.01:
            continue
        g = lambda c: f"{r[idx[c]]} {units[idx[c]]}" if c in idx else ''
        out.append(f"| `{name}` | {r[idx['launch__grid_size']]} | {g(cols[0])} | {g(cols[1])} | {g(cols[2])} | "
                   f"{float(r[idx[cols[3]]]):.1f} | {float(r[idx[cols[4]]]):.1f} | {float(r[idx[cols[5]]]):.1f} | "
                   f"{r[idx[cols[6]]]} |")
notes = os.path.join(P, f'{tag}_notes.md')     # hand-written sections (other configs, probes) travel with the summary
if os.path.exists(notes):
    out.append('\n' + open(notes).read().rstrip())
open(os.path.join(P, f'{tag}_summary.md'), 'w').write('\n'.join(out) + '\n')
print('\n'.join(out)[:3000])
