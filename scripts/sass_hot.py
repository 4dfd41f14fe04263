"""Summarise an `ncu -i X.ncu-rep --page source --csv --print-source sass [--launch-skip N --launch-count 1]` dump:
hottest SASS address ranges by executed warp-instructions and stall samples.
usage: python scripts/sass_hot.py dump.csv [min_share]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Address" in r and "Source" in r)
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
iA, iS, iE, iT, iSm = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "Avg. Threads Executed", "# Samples"))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
tot = sum(int(r[iE]) for r in data)
totS = sum(int(r[iSm]) for r in data)
print(f"warp-instructions {tot:.4g}, stall samples {totS}, SASS lines {len(data)}")
base = int(data[0][iA], 16)
groups, cur = [], []
for r in data:
    n = int(r[iE])
    if cur and abs(n - int(cur[-1][iE])) > 0.03 * max(n, int(cur[-1][iE]), 1):
        groups.append(cur)
        cur = []
    cur.append(r)
groups.append(cur)
for g in groups:
    n = sum(int(r[iE]) for r in g)
    s = sum(int(r[iSm]) for r in g)
    if n / tot < thr and s / totS < thr:
        continue
    ops = {}
    for r in g:
        t = r[iS].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    print(f"{int(g[0][iA], 16) - base:5x}-{int(g[-1][iA], 16) - base:5x} n={len(g):3d} exec/ins={int(g[0][iE]):10d} share={n / tot:6.2%} "
          f"samples={s / totS:6.2%} thr={sum(float(r[iT]) for r in g) / len(g):5.1f} | " + " ".join(f"{o}x{c}" for o, c in sorted(ops.items(), key=lambda kv: -kv[1])[:7]))
