"""Maps an `ncu --page source --csv` SASS dump onto CUDA source lines using nvdisasm -g line markers.

usage: ncu_hot_lines.py <sass.csv> <kernel-substring> <nvdisasm -g -c output> [top]
Aggregates warp-stall samples and executed instructions per (file, line) for the first launch.
"""
import csv, re, sys, collections

csv_path, kern, asm_path = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# ---- 1. instruction index -> (file, line) from nvdisasm
lines_of = []
cur = None
infunc = False
for l in open(asm_path, errors="replace"):
    if l.startswith(".text.") or l.lstrip().startswith(".section\t.text."):
        infunc = kern in l
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines_of.append(cur)

# ---- 2. samples per instruction (first launch of the kernel)
rows = list(csv.reader(open(csv_path)))
hdr = None
data = []
started = False
for r in rows:
    if r and r[0] == "Kernel Name":
        if started:
            break
        continue
    if r and r[0] == "Address":
        hdr = r
        started = True
        continue
    if started:
        data.append(r)
ci = {n: i for i, n in enumerate(hdr)}
S, X, T = ci["# Samples"], ci["Instructions Executed"], ci["Thread Instructions Executed"]
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
n = min(len(data), len(lines_of))
for k in range(n):
    r = data[k]
    key = lines_of[k]
    v = (int(r[S] or 0), int(r[X] or 0), int(r[T] or 0))
    for j in range(3):
        agg[key][j] += v[j]
        tot[j] += v[j]
print(f"instructions: csv {len(data)} asm {len(lines_of)}; samples {tot[0]} warp-inst {tot[1]} thread-inst {tot[2]} (avg threads {tot[2]/max(tot[1],1):.1f})")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(key):40s} samples {v[0]:7d} ({v[0]/tot[0]*100:5.1f}%)  inst {v[1]/tot[1]*100:5.1f}%  avg thr {v[2]/max(v[1],1):5.1f}")
