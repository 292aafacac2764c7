"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel
(B200_PROFILING.md "What proves a Blackwell-native kernel"): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM,
TMA -> UTMALDG/UTMASTG/UBLKCP, tcgen05.commit -> UTCBAR, legacy mma.sync -> HMMA.  Writes profiles/<tag>_sass_summary.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'content-aware-gan-compression_b200', 'lib', 'libcagc_b200.so')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r2'
sass = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
pat = re.compile(r'\b(UTC[A-Z]*MMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCATOMSWS|HMMA|HGMMA|SYNCS|ELECT|FFMA2|LDGSTS|R2UR)\b')
counts, cur, order = collections.defaultdict(collections.Counter), None, []
ninstr = collections.Counter()
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip().split('(')[0]
        order.append(cur)
        continue
    if cur and re.search(r'/\*[0-9a-f]{4,}\*/\s+\S', line):
        ninstr[cur] += 1
        for mm in pat.findall(line):
            counts[cur][mm] += 1
cols = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'SYNCS', 'ELECT', 'FFMA2', 'HMMA']
out = [f'# SASS mnemonic counts per kernel, {os.path.relpath(LIB, ROOT)} (cuobjdump -sass), sm_100a', '',
       f'{"kernel":64s} {"instr":>6s} ' + ' '.join(f'{c:>7s}' for c in cols)]
for k in order:
    c = counts[k]
    other = sum(v for kk, v in c.items() if kk.startswith('UTC') and kk.endswith('MMA'))
    row = [other if col == 'UTCHMMA' else c.get(col, 0) for col in cols]
    out.append(f'{k[:64]:64s} {ninstr[k]:6d} ' + ' '.join(f'{v:7d}' for v in row))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
out += ['', 'totals: ' + ', '.join(f'{k} {v}' for k, v in sorted(tot.items()))]
out += ['', 'UTC*MMA = tcgen05.mma (kind::tf32), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA load), UTCBAR = tcgen05.commit,',
        'SYNCS = mbarrier ops, FFMA2 = packed fma.rn.f32x2 (FIR), HMMA would be a legacy mma.sync path (none).',
        'No UTMASTG: epilogues store through registers (transposed in shared memory for N >= 128, see DESIGN.md section 4).']
path = os.path.join(ROOT, 'profiles', f'{tag}_sass_summary.txt')
open(path, 'w').write('\n'.join(out) + '\n')
print('\n'.join(out))
