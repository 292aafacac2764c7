"""`ncu -i X.ncu-rep --page raw --csv` -> one compact CSV row per launch (time, DRAM traffic, pipe utilisation)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [('time', 'gpu__time_duration.sum'), ('dram_rd', 'dram__bytes_read.sum'), ('dram_wr', 'dram__bytes_write.sum'),
        ('dram%', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        ('tensor%', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        ('lts%', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
        ('lsu%', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'),
        ('regs', 'launch__registers_per_thread'), ('warps_active%', 'sm__warps_active.avg.pct_of_peak_sustained_active')]
cols = [(n, c) for n, c in cols if c in idx]
print('kernel,grid,block,' + ','.join(f'{n}({units[idx[c]]})' if units[idx[c]] else n for n, c in cols))
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0]
    print(','.join([name.replace(',', ';'), r[idx['launch__grid_size']], r[idx['launch__block_size']]] +
                   [r[idx[c]].replace(',', '') for _, c in cols]))
