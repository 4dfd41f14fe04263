#!/bin/bash
# both arms of the bench with the shared `config` dict
O=gpurun_out
mkdir -p $O
python bench.py --impl reference --steps 1 --warmup 1 > $O/r02_bench_reference.json 2>/dev/null
( time python bench.py > $O/r02_bench_default_1gpu.json 2> $O/r02_bench_default.err ) 2>&1 | tail -4
python - <<'PY'
import json
l = json.loads(open('gpurun_out/r02_bench_default_1gpu.json').read().strip().splitlines()[-1])
r = json.loads(open('gpurun_out/r02_bench_reference.json').read().strip().splitlines()[-1])
print('same config:', l['config'] == r['config'], '| same metric/unit:', (l['metric'], l['unit'], l['higher_is_better']) == (r['metric'], r['unit'], r['higher_is_better']))
print('value %.4g e2e %.4g ms %.1f launches %s frac %.3f traffic %s | reference %.4g | e2e ratio %.0f' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches'], l['roofline']['frac'], l['roofline']['traffic'], r['value'], l['e2e']['value'] / r['value']))
for k in ('full_table', 'compat', 'non_invariant', 'gradient_scales', 'other_configs', 'north_star'):
    v = l.get(k)
    if isinstance(v, dict) and 'value' in v: print(k, '%.4g' % v['value'], v.get('error', ''))
    else: print(k, {kk: ('%.4g' % vv['value'] if isinstance(vv, dict) and 'value' in vv else vv) for kk, vv in (v or {}).items()} if isinstance(v, dict) else v)
PY
