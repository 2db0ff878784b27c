"""Count local-memory ops inside the tensor-layer epilogue loops of the <2,2,pure,tanh> instantiation."""
import re, subprocess, sys
obj = sys.argv[1]
out = subprocess.run("cuobjdump -elf %s" % obj, shell=True, capture_output=True, text=True).stdout
syms = {}
for line in out.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+0x2\s+\S+\s+0x[0-9a-f]+\s+(\S+)", line)
    if m: syms[m.group(3).split('$')[-1]] = (int(m.group(1), 16), int(m.group(2), 16))
sass = subprocess.run("cuobjdump -sass %s" % obj, shell=True, capture_output=True, text=True).stdout
ins = []
for l in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
for key in ("net_forwardILi2ELi2ELb1ELi1E", "net_backwardILi2ELi2ELb1ELi1E"):
    for name, (lo, sz) in syms.items():
        if key in name:
            f = [t for a, t in ins if lo <= a < lo + sz]
            ld = [i for i, t in enumerate(f) if "LDTM" in t]
            regs = set()
            for t in f:
                for r in re.findall(r"\bR(\d+)\b", t): regs.add(int(r))
            # loop body ~ from first LDTM of a cluster to the next backward BRA
            print(key, "instrs", len(f), "max reg", max(regs), "total LDL/STL", sum(("LDL" in t or "STL" in t) for t in f))
            clusters = []
            for i in ld:
                if not clusters or i - clusters[-1][-1] > 40: clusters.append([i])
                else: clusters[-1].append(i)
            for c in clusters:
                a = c[0]
                b = next((j for j in range(c[-1], len(f)) if f[j].startswith("@") and "BRA" in f[j] or f[j].startswith("BRA")), len(f) - 1)
                body = f[a:b + 1]
                print("   loop@%d len %d  LDTM %d  LDL %d STL %d MUFU %d" % (a, len(body), len(c), sum("LDL" in t for t in body),
                      sum("STL" in t for t in body), sum("MUFU" in t for t in body)))
