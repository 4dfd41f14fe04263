#!/bin/bash
# state check after the container was re-created: all GPU tests, the default bench line, the reference arm
O=gpurun_out
mkdir -p $O; rm -f $O/parity_report.txt
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -12 ) 2>&1 | tee $O/r02s_pytest_gpu.log
cp $O/parity_report.txt $O/r02s_parity_report.txt 2>/dev/null
( time python bench.py > $O/r02s_bench_default.json 2> $O/r02s_bench_default.err ) 2>&1 | tail -4
python - <<'PY'
import json
l = json.loads(open('gpurun_out/r02s_bench_default.json').read().strip().splitlines()[-1])
print('value %.4g e2e %.4g ms %.1f launches %s frac %.3f' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches'], l['roofline']['frac']))
for k in ('full_table', 'compat', 'non_invariant', 'north_star'):
    if k in l: print(k, json.dumps(l[k])[:600])
PY
