"""Aggregate an ncu launch list (gpu__time_duration + dram bytes per launch) by kernel name."""
import collections, csv, sys
path = sys.argv[1]
rows = list(csv.DictReader([l for l in open(path) if not l.startswith('==')]))
per = collections.defaultdict(dict)
for r in rows:
    per[r['ID']]['name'] = r['Kernel Name']
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    m = r['Metric Name']
    if m == 'gpu__time_duration.sum':
        v = v / 1e3 if u in ('nsecond', 'ns') else v * 1e3 if u in ('msecond', 'ms') else v   # -> us
    else:
        v = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1) * v
    per[r['ID']][m] = v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in per.values():
    a = agg[d['name'].split('(')[0][:70]]
    a[0] += 1; a[1] += d.get('gpu__time_duration.sum', 0); a[2] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
tot = sum(a[1] for a in agg.values()); totb = sum(a[2] for a in agg.values())
print(f'{len(per)} launches, {tot/1e3:.2f} ms serialized, {totb/1e9:.2f} GB DRAM')
print(f'{"kernel":70s} {"n":>4s} {"us":>9s} {"share":>6s} {"GB":>7s} {"GB/s":>7s}')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f'{k:70s} {a[0]:4d} {a[1]:9.1f} {100*a[1]/tot:5.1f}% {a[2]/1e9:7.3f} {a[2]/max(a[1],1e-9)/1e3:7.0f}')
