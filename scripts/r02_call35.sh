#!/bin/bash
# full set + hottest SASS ranges of the shipped MULTI kernel (C3, 2e6 spins)
O=gpurun_out
mkdir -p $O
ncu --set full --import-source on --clock-control none -k regex:walk_fast -c 2 -o $O/r02_full_c3 -f python bench.py --workload c3 --spins 2000000 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras > $O/r02_full_c3.log 2>&1
ncu -i $O/r02_full_c3.ncu-rep --page details --csv > $O/r02_full_c3_details.csv 2>/dev/null
ncu -i $O/r02_full_c3.ncu-rep --page source --csv --print-source sass --launch-skip 1 --launch-count 1 > $O/r02_full_c3_source.csv 2>/dev/null
rm -f $O/r02_full_c3.ncu-rep
python scripts/sass_hot.py $O/r02_full_c3_source.csv 0.01 | awk '!seen[$0]++' > $O/r02_sass_hot_c3.txt
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_full_c3_source.csv')))
hdr=next(r for r in rows if "Address" in r and "Source" in r)
data=[r for r in rows if len(r)==len(hdr) and r[0].startswith("0x")]
iA,iS,iE,iSm=(hdr.index(k) for k in ("Address","Source","Instructions Executed","# Samples"))
tot=sum(int(r[iE]) for r in data); totS=sum(int(r[iSm]) for r in data)
# the out-of-line callee starts after the kernel's last EXIT
last_exit=max(i for i,r in enumerate(data) if r[iS].strip().startswith("EXIT") or " EXIT" in r[iS][:12])
k=sum(int(r[iE]) for r in data[:last_exit+1]); ks=sum(int(r[iSm]) for r in data[:last_exit+1])
print("kernel body: %.1f %% of the warp-instructions, %.1f %% of the stall samples; out-of-line functions (events, start): %.1f %% / %.1f %%" % (100*k/tot, 100*ks/totS, 100-100*k/tot, 100-100*ks/totS))
PY
rm -f $O/r02_full_c3_source.csv
