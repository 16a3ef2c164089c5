"""Summarise an ncu source-page CSV (ncu -i X.ncu-rep --page source --print-source sass --csv): stall samples by reason,
by opcode, the top SASS lines, and shared-memory wavefront excess.  Usage: python tools/ncu_src.py file.csv [ntop]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) >= len(hdr) - 2]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25


def f(r, name):
    try:
        return float(r[col[name]])
    except (ValueError, IndexError):
        return 0.0


tot = sum(f(r, "# Samples") for r in data)
inst = sum(f(r, "Instructions Executed") for r in data)
print("total samples", tot, "warp instructions", inst)
reasons = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
rs = Counter({n: sum(f(r, n) for r in data) for n in reasons})
print("stall reasons:", ", ".join(f"{n[6:]} {100 * v / tot:.1f}%" for n, v in rs.most_common(9)))
byop_s, byop_i = Counter(), Counter()
for r in data:
    op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[col["Source"]].split()[1]
    op = op.split(".")[0]
    byop_s[op] += f(r, "# Samples")
    byop_i[op] += f(r, "Instructions Executed")
print("by opcode (samples%, instr%):", ", ".join(f"{o} {100 * s / tot:.1f}/{100 * byop_i[o] / inst:.1f}" for o, s in byop_s.most_common(14)))
wf, wfi = sum(f(r, "L1 Wavefronts Shared") for r in data), sum(f(r, "L1 Wavefronts Shared Ideal") for r in data)
print(f"shared wavefronts {wf:.3g} ideal {wfi:.3g}")
print("top lines:")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:ntop]:
    top = max(reasons, key=lambda n: f(r, n))
    print(f"  {r[col['Address']][-5:]} {100 * f(r, '# Samples') / tot:5.2f}% inst {100 * f(r, 'Instructions Executed') / inst:5.2f}% {top[6:]:>12} wf {f(r, 'L1 Wavefronts Shared'):.3g}/{f(r, 'L1 Wavefronts Shared Ideal'):.3g}  {r[col['Source']][:70]}")

# loops: a backward branch closes the region [target, branch]
addr = [int(r[col["Address"]], 16) for r in data]
base = addr[0]
print("loops (innermost regions closed by a backward branch):")
loops = []
for i, r in enumerate(data):
    src = r[col["Source"]]
    if "BRA" in src and "0x" in src:
        try:
            tgt = int(src.split("0x")[-1].split()[0].rstrip(";"), 16)
        except ValueError:
            continue
        if tgt < addr[i]:
            j = next((k for k, a in enumerate(addr) if a >= tgt), None)
            if j is not None and i - j < 400:
                loops.append((j, i))
for (j, i) in sorted(loops, key=lambda ji: -sum(f(r, "# Samples") for r in data[ji[0]:ji[1] + 1]))[:14]:
    seg = data[j:i + 1]
    s = sum(f(r, "# Samples") for r in seg)
    rc = Counter({n: sum(f(r, n) for r in seg) for n in reasons})
    ops = Counter()
    for r in seg:
        t = r[col["Source"]].split()
        op = (t[1] if t and t[0].startswith("@") else (t[0] if t else "?")).split(".")[0]
        ops[op] += 1
    execs = max(f(r, "Instructions Executed") for r in seg)
    print(f"  +{addr[j] - base:05x}..+{addr[i] - base:05x} n={len(seg):4d} samples {100 * s / tot:5.2f}% instr {100 * sum(f(r, 'Instructions Executed') for r in seg) / inst:5.2f}% "
          f"trips {execs:.3g} | " + ", ".join(f"{n[6:]} {100 * v / max(s, 1):.0f}%" for n, v in rc.most_common(5)) + " | " +
          ", ".join(f"{o}x{c}" for o, c in ops.most_common(5)))
