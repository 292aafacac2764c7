"""Development aid: print the per-layer records of a `bench.py --all-layers` JSON line, heaviest first."""
import json
import sys
d = json.load(open(sys.argv[1]))
pat = sys.argv[2] if len(sys.argv) > 2 else ''
rows = sorted(d['all_layers'].items(), key=lambda kv: -kv[1]['avg_launch_us'] * kv[1]['launches_per_step'])
print(round(d['ms_per_step'], 3), {k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})
for k, v in rows:
    if pat in k:
        print(f"{v['avg_launch_us'] * v['launches_per_step']:8.1f}  {k}  n={v['launches_per_step']:.0f} avg={v['avg_launch_us']:.1f} "
              f"tf={v.get('tflops') or 0:.0f} gbs={v.get('gbs') or 0:.0f}")
