"""Summarise an `ncu --page source --csv --print-source sass` dump: hottest SASS ranges by executed instructions.
usage: python scripts/sass_hot.py dump.csv [min_exec_frac]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iA, iS, iE, iT, iSm = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tot = sum(int(r[iE]) for r in data)
totS = sum(int(r[iSm]) for r in data)
print(f"total warp-instructions {tot:.4g}, samples {totS}")
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
# print every instruction with exec count, grouped into runs of equal count
prev = None
run = []
def flush():
    if not run: return
    n = int(run[0][iE]); k = len(run)
    if n * k / tot >= thr:
        smp = sum(int(r[iSm]) for r in run)
        st = {}
        for r in run:
            for c in stall_cols:
                v = int(r[c] or 0)
                if v: st[hdr[c]] = st.get(hdr[c], 0) + v
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        ops = {}
        for r in run:
            op = r[iS].split()[0] if not r[iS].strip().startswith("@") else r[iS].split()[1]
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        opss = " ".join(f"{o}x{c}" for o, c in sorted(ops.items(), key=lambda kv: -kv[1])[:8])
        print(f"{run[0][iA][-5:]} n={k:4d} exec/instr={n:11d} share={n*k/tot:6.2%} samples={smp/totS:6.2%} thr={float(run[0][iT]):5.1f} | {opss} | {top}")
for r in data:
    key = r[iE]
    if prev is not None and key != prev:
        flush(); run = []
    run.append(r); prev = key
flush()
