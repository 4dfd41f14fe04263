#!/bin/bash
# after the bounded magnetisation buffer of the MULTI kernels: tests, re-stamped few-metric captures (scripts/r02_profile.sh without the full sets), bench line
O=gpurun_out
mkdir -p $O; rm -f $O/parity_report.txt
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -12 ) 2>&1 | tee $O/r02_gpu_tests.log
bash scripts/r02_profile.sh quick > $O/r02_profile_quick.log 2>&1; tail -4 $O/r02_profile_quick.log
cp $O/traffic.json profiles/traffic.json
( time python bench.py > $O/r02_bench_default_1gpu.json 2> $O/r02_bench_default.err ) 2>&1 | tail -4
python - <<'PY'
import json
l = json.loads(open('gpurun_out/r02_bench_default_1gpu.json').read().strip().splitlines()[-1])
print('value %.4g e2e %.4g ms %.1f launches %s frac %.3f traffic %s' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches'], l['roofline']['frac'], l['roofline']['traffic']))
print('traffic_source', l['roofline']['traffic_source'][:100])
for k in ('full_table', 'compat', 'non_invariant', 'gradient_scales', 'other_configs', 'north_star'):
    v = l.get(k)
    if isinstance(v, dict) and 'value' in v: print(k, '%.4g' % v['value'], v.get('error', ''))
    else: print(k, {kk: ('%.4g' % vv['value'] if isinstance(vv, dict) and 'value' in vv else vv) for kk, vv in (v or {}).items()} if isinstance(v, dict) else v)
PY
