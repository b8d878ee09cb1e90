"""Summarise an ncu capture of k_value_tower: python tools/tower_ncu_summary.py <rep> <leaves per launch> <flops per leaf>."""
import csv, subprocess, sys
rep, leaves, flops_leaf = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, v = rows[0], rows[1], rows[2]
d = {h: (x, u) for h, u, x in zip(hdr, units, v)}
print("kernel:", d['Kernel Name'][0][:110])
print(f"leaves per launch: {leaves}")
for k in ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'launch__grid_size', 'launch__cluster_size' if 'launch__cluster_size' in d else 'launch__block_size', 'launch__block_size', 'launch__registers_per_thread',
          'launch__shared_mem_per_block_dynamic', 'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
          'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
          'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
          'smsp__issue_active.avg.pct_of_peak_sustained_active']:
    if k in d:
        print(f"  {k}: {d[k][0]} {d[k][1]}")
ms = float(d['gpu__time_duration.sum'][0])
print(f"  => {leaves * flops_leaf / ms / 1e9:.0f} TFLOP/s algorithmic at this (profiler, unthrottled-clock) launch")
def tobytes(x,u):
    x=float(x); return x*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[u]
tr = tobytes(*d['dram__bytes_read.sum'])+tobytes(*d['dram__bytes_write.sum'])
print(f"  => DRAM traffic {tr/1e6:.1f} MB per launch = {tr/leaves:.1f} B per leaf")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(h)]
ix = {k: h.index(k) for k in h}
from collections import Counter
c = Counter()
for r in data:
    t = r[ix['Source']].strip().split()
    if not t: continue
    op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
    c[op.split('.')[0]] += float(r[ix['Instructions Executed']] or 0)
print("  SASS evidence (instructions executed):", {k: int(c[k]) for k in ('UTCHMMA', 'UTCBAR', 'LDTM', 'UBLKCP', 'SYNCS') if k in c})
