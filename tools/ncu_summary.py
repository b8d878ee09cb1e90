"""Summarise an .ncu-rep (raw + source pages) into a few lines (development aid; needs ncu on PATH)."""
import csv, subprocess, sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__occupancy_limit_registers']
for r in rows[2:3]:
    print("kernel:", r[hdr.index('Kernel Name')][:70])
    for k in keys:
        if k in hdr:
            print(f"  {k}: {r[hdr.index(k)]} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(h)]
ix = {k: h.index(k) for k in h}
f = lambda r, k: float(r[ix[k]]) if r[ix[k]] not in ('', 'N/A') else 0.0
st = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
tot = {k: sum(f(r, k) for r in data) for k in st}
s = sum(tot.values()) or 1
print("  stalls:", {k[6:]: round(100 * v / s, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]})
c = Counter()
for r in data:
    t = r[ix['Source']].strip().split()
    if not t:
        continue
    op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
    c[op.split('.')[0]] += f(r, 'Instructions Executed')
t = sum(c.values()) or 1
print("  opcodes:", [(op, round(100 * n / t, 1)) for op, n in c.most_common(14)])
