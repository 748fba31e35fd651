"""top stall-sampled SASS instructions of an `ncu --page source --csv` export: python scripts/ncu_src_top.py file.csv [N]"""
import csv
import sys
from collections import Counter

f = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(f)))[2:]
tot = sum(int(r[2] or 0) for r in rows)
inst = sum(int(r[5] or 0) for r in rows)
print("instructions", len(rows), "samples", tot, "warp-inst executed", inst)
ops = Counter()
for r in rows:
    op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
    ops[op.split(".")[0]] += int(r[5] or 0)
print("executed by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / inst) for k, v in ops.most_common(18)))
top = sorted(range(len(rows)), key=lambda i: -int(rows[i][2] or 0))[:n]
for i in sorted(top):
    r = rows[i]
    print("%5d %6.2f%% exec %10s thr/warp %5s  %s" % (i, 100.0 * int(r[2] or 0) / tot, r[5], r[8], r[1].strip()[:110]))
