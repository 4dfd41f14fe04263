#!/bin/bash
# one walk for all gradient / phase-cycling scales: tests, then C3 / C3r with and without it, C2 unchanged?
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_fast_parity_gpu.py -m gpu -q -x -k "one_walk or shared_and_private or c3 or pgse or random_cases or fast_mode" 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -25 | tee $O/r02t_pytest.log
for wl in c3 c3r; do
  for v in "" "SWK_NO_ONEWALK=1"; do
    env $v python bench.py --workload $wl --spins 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>$O/r02t_err.log | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('$wl $v value %.4g e2e %.4g ms %.2f launches %s' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches']))
" | tee -a $O/r02t_c3.log
  done
done
python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/r02t_bench_c3.json 2>>$O/r02t_err.log
python bench.py --workload c3r --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/r02t_bench_c3r.json 2>>$O/r02t_err.log
python bench.py --no-cpu-baseline --no-extras --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('c2 value %.4g e2e %.4g ms %.1f launches %s' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches']))
" | tee -a $O/r02t_c3.log
python -c "
import json
for w in ('c3','c3r'):
    l=json.loads(open('gpurun_out/r02t_bench_%s.json'%w).read().strip().splitlines()[-1])
    print(w, 'full size value %.4g e2e %.4g ms %.2f' % (l['value'], l['e2e']['value'], l['ms_per_step']))
" | tee -a $O/r02t_c3.log
tail -5 $O/r02t_err.log
