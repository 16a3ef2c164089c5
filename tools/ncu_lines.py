"""Per-CUDA-source-line stall samples from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
Usage: python tools/ncu_lines.py file.csv [ntop] [barrier-exclude 0/1]"""
import csv
import sys
from collections import Counter, defaultdict

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, hdr, col = None, None, None
samples, bar, instr, text, reasons = Counter(), Counter(), Counter(), {}, defaultdict(Counter)
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        col = {}
        for i, n in enumerate(hdr):
            col.setdefault(n, i)
        continue
    if hdr is None or len(r) < 8:
        continue
    if r[0] != "":
        cur = (fname, int(r[0]))
        text[cur] = r[1].strip()
        continue
    if cur is None:
        continue
    try:
        s = float(r[col["# Samples"]])
    except ValueError:
        continue
    samples[cur] += s
    instr[cur] += float(r[col["Instructions Executed"]])
    for n in hdr:
        if n.startswith("stall_") and "Not Issued" not in n:
            reasons[cur][n[6:]] += float(r[col[n]])
tot = sum(samples.values())
ti = sum(instr.values())
print("total samples", tot)
for k, s in samples.most_common(ntop):
    rs = ", ".join(f"{n} {100 * v / max(s, 1):.0f}%" for n, v in reasons[k].most_common(3))
    print(f"{100 * s / tot:5.2f}% instr {100 * instr[k] / ti:5.2f}%  {k[0]}:{k[1]:<4d} {text[k][:80]:80s} | {rs}")
