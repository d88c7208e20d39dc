"""Per-launch summary table of an .ncu-rep (ncu --set full): python profiles/ncu_table.py REP"""
import csv, re, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h = r[0]
keys = [('us', 'gpu__time_duration.sum'), ('tensor pipe %', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active'),
        ('fma pipe %', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'),
        ('issue %', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
        ('L2 %', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'), ('warps %', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
        ('regs', 'launch__registers_per_thread'), ('dram rd MB', 'dram__bytes_read.sum'), ('dram wr MB', 'dram__bytes_write.sum'),
        ('M warp-instr', 'smsp__inst_executed.sum')]
keys = [(k, c) for k, c in keys if c in h]
print('| kernel | ' + ' | '.join(k for k, _ in keys) + ' | dram GB/s |')
print('|---|' + '---:|' * (len(keys) + 1))
for row in r[2:]:
    name = re.sub(r'\(.*', '', row[h.index('Kernel Name')]).replace('void ', '').replace('advmil::', '').replace('(int)', '').replace('(bool)', '')[:64]
    vals = []
    for k, c in keys:
        i = h.index(c)
        try:
            v = float(row[i].replace(',', '') or 0)
        except ValueError:
            v = float('nan')
        u = r[1][i]
        if 'MB' in k:
            v *= {'Gbyte': 1e3, 'Mbyte': 1, 'Kbyte': 1e-3, 'byte': 1e-6}.get(u, 1)
        if k == 'us':
            v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(u, 1)
        if k == 'M warp-instr':
            v /= 1e6
        vals.append(f'{v:.1f}')
    d = dict(zip([k for k, _ in keys], vals))
    gbs = (float(d['dram rd MB']) + float(d['dram wr MB'])) / float(d['us']) * 1e3 if 'dram rd MB' in d else float('nan')
    print('| `' + name + '` | ' + ' | '.join(vals) + f' | {gbs:.0f} |')
