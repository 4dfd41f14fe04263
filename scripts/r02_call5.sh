#!/bin/bash
O=gpurun_out
mkdir -p $O; rm -f $O/parity_report.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -30 | tee $O/r02e_pytest_gpu.log
( time python bench.py > $O/r02e_bench_c2.json 2> $O/r02e_bench_c2.err ) 2>&1 | tail -3
tail -5 $O/r02e_bench_c2.err
python -c "
import json
l=json.load(open('$O/r02e_bench_c2.json'))
print('value %.4g e2e %.4g ms %.1f launches %s' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches']))
print('roofline', {k:l['roofline'][k] for k in ('achieved','frac','bytes_per_spin_step','traffic','traffic_source')})
print('cpu', l.get('cpu_baseline',{}).get('value'), 'refcuda', (l.get('reference_cuda') or {}).get('value'))
for k in ('full_table','compat','non_invariant','north_star'):
    r=l.get(k,{})
    print(k, r.get('value'), r.get('ms_per_step'), r.get('error'), (r.get('roofline') or {}).get('gather'))
print('groups', l.get('scale_groups'))
print('ns full', (l.get('north_star') or {}).get('full_table'))
print('ns cpu', (l.get('north_star') or {}).get('cpu_baseline'))
"
PROBE_FLAGS=512 python scripts/group_probe.py 2000000 c2 "full-default" "full-hybrid:SWK_SHARE_SIGMA=1.25" 2>&1 | tee $O/r02e_groups_c2_full.log
ls -la $O | tail -4
