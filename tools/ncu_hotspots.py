"""Where do a kernel's executed instructions and stall samples go, by source function and line?
Joins the SASS rows of `ncu --page source --csv` (one per instruction, in address order) with the line table of the
cubin (`nvdisasm -gi`), then attributes every instruction to the function that contains its innermost source line and
to the outermost inlined call site.  Development aid; needs ncu, cuobjdump and nvdisasm on PATH.

  python tools/ncu_hotspots.py gpurun_out/x.ncu-rep zeroclone_b200/libzc_b200.so <mangled-kernel-substring> [top]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()

# ---- line table of the kernel: offset -> [(file, line), ...] innermost first
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
table, cur = {}, []
pat_file = re.compile(r'//## File "([^"]+)", line (\d+)(.*)')
pat_inl = re.compile(r'inlined at "([^"]+)", line (\d+)')
pat_ins = re.compile(r'^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);')
group_open = False
for l in dis[start + 1:]:
    if l.startswith("//-------") or l.startswith(".text."):
        break
    m = pat_file.search(l)
    if m:       # consecutive "File" lines form one chain: innermost first, then each call site outwards
        if not group_open:
            cur, group_open = [], True
        cur.append((m.group(1), int(m.group(2))))
        continue
    m = pat_ins.match(l)
    if m:
        group_open = False
        table[int(m.group(1), 16)] = (list(cur), m.group(2).strip())

# ---- enclosing function of a source line (crude: the last definition-looking line above it)
fn_cache = {}
pat_def = re.compile(r'^\s*(?:template\s*<[^>]*>\s*)?(?:ZC_HD|ZC_D|__device__|__global__|static|inline|__forceinline__|__noinline__|constexpr|\s)*[\w:<>\*&, ]+?\b(\w+)\s*\([^;]*$')


def enclosing(path, line):
    if path not in fn_cache:
        names = []
        try:
            src = open(path).read().splitlines()
        except OSError:
            src = []
        cur_name = os.path.basename(path)
        head = re.compile(r'^\s*(?:ZC_HD|ZC_D|__device__|__global__|__host__|static\s+(?:ZC_HD|ZC_D|__device__)|template\s*<)')
        for i, s in enumerate(src, 1):
            st = s.rstrip()
            if head.match(s) and "(" in s and not st.endswith(";"):
                body = s.replace("__launch_bounds__(", "LB_").replace("__align__(", "AL_")
                m = re.search(r'(\w+)\s*\(', body[re.search(r'(ZC_HD|ZC_D|__device__|__global__)', body).end():] if re.search(r'(ZC_HD|ZC_D|__device__|__global__)', body) else body)
                if m and m.group(1) not in ("if", "for", "while", "switch", "return", "sizeof"):
                    cur_name = m.group(1)
            names.append(cur_name)
        fn_cache[path] = names
    names = fn_cache[path]
    return names[line - 1] if 0 < line <= len(names) else os.path.basename(path)


# ---- ncu rows
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ix = {k: i for i, k in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) == len(h)]
base = int(data[0][0], 16)


def f(r, k):
    v = r[ix[k]]
    return float(v) if v not in ("", "N/A", "-") else 0.0


by_fn, by_site, by_line = defaultdict(Counter), defaultdict(Counter), defaultdict(Counter)
tot = Counter()
hot_floor = 0.01 * max(f(r, "Instructions Executed") for r in data)     # "hot" = executed at least 1 % as often as the hottest instruction
for r in data:
    off = int(r[0], 16) - base
    stack, _ = table.get(off, ([("?", 0)], ""))
    inner = stack[0]
    fn = enclosing(*inner) if inner[0] != "?" else "?"
    # the chain of enclosing functions, outermost first (kernel > ... > innermost), cut to three levels
    chain = []
    for pth, ln in reversed(stack):
        nm = enclosing(pth, ln) if pth != "?" else "?"
        if not chain or chain[-1] != nm:
            chain.append(nm)
    site = " > ".join(chain[1:4]) if len(chain) > 1 else chain[0]
    m = {"inst": f(r, "Instructions Executed"), "thread_inst": f(r, "Thread Instructions Executed"), "samples": f(r, "# Samples"),
         "no_inst": f(r, "stall_no_inst"), "long_sb": f(r, "stall_long_sb"), "wait": f(r, "stall_wait"), "size": 1,
         "hot": 1 if f(r, "Instructions Executed") >= hot_floor else 0}
    for k, v in m.items():
        by_fn[fn][k] += v
        by_site[site][k] += v
        by_line["%s:%d" % (os.path.basename(inner[0]), inner[1])][k] += v
        tot[k] += v


def show(title, d, n):
    print(f"\n== {title} (share of executed warp instructions | lanes active | share of stall samples | no_inst share of its samples | static instructions | hot KB)")
    for name, c in sorted(d.items(), key=lambda kv: -kv[1]["inst"])[:n]:
        lanes = c["thread_inst"] / c["inst"] if c["inst"] else 0
        print(f"  {name:46s} {100 * c['inst'] / tot['inst']:5.1f} %  {lanes:4.1f}  {100 * c['samples'] / max(1, tot['samples']):5.1f} %  "
              f"{100 * c['no_inst'] / max(1, c['samples']):5.1f} %  {int(c['size']):5d}  {c['hot'] * 16 / 1024:5.1f}")


print(f"kernel {kern}: {int(tot['size'])} static instructions ({int(tot['size']) * 16 / 1024:.0f} KB), {tot['inst']:.3e} executed, "
      f"{tot['thread_inst'] / tot['inst']:.1f} lanes, no_inst {100 * tot['no_inst'] / max(1, tot['samples']):.1f} % of samples, hot code {tot['hot'] * 16 / 1024:.1f} KB")
show("by function containing the instruction's source line", by_fn, top)
show("by call chain below the kernel (three levels)", by_site, top)
show("by source line", by_line, top)
