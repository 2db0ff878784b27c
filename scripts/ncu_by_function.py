"""Attribute an ncu SASS source page (csv) to the sub-functions of a kernel using the cubin symbol table.
usage: ncu -i rep --page source --csv > src.csv ; python scripts/ncu_by_function.py src.csv build/obj.o"""
import csv, re, subprocess, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
H = rows[1]; idx = {h: i for i, h in enumerate(H)}
S = idx['# Samples']; IE = idx['Instructions Executed']; A = idx['Address']
data = []
for r in rows[2:]:
    try:
        data.append([int(r[A], 16), int(float(r[S] or 0)), int(float(r[IE] or 0)), r[idx['Source']]])
    except Exception:
        pass
base = data[0][0]
out = subprocess.run("cuobjdump -elf %s" % sys.argv[2], shell=True, capture_output=True, text=True).stdout
syms = []
for line in out.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+0x2\s+\S+\s+0x[0-9a-f]+\s+(\S+)", line)
    if m:
        syms.append((int(m.group(1), 16), int(m.group(2), 16), m.group(3)))
syms.sort()
tot_s = sum(d[1] for d in data); tot_i = sum(d[2] for d in data)
agg = collections.defaultdict(lambda: [0, 0])
for a, s, i, src in data:
    off = a - base
    name = 'kernel_main'
    for o, sz, n in syms:
        if o <= off < o + sz:
            name = n.split('$')[-1][9:64]
            break
    agg[name][0] += s; agg[name][1] += i
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    if v[0] > 0:
        print("%-58s samples %5.1f%%  instr %5.1f%% (%d)" % (k, 100 * v[0] / tot_s, 100 * v[1] / tot_i, v[1]))
