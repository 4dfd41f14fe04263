#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_fast_parity_gpu.py -m gpu -q -x -k "one_walk or c3 or pgse or rebinned" 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -25 | tee $O/r02_pytest_multi2.log
for wl in c3 c3r; do
  python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-e2e 2>$O/r02w_err.log | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('$wl value %.4g ms %.2f launches %s' % (l['value'], l['ms_per_step'], l['gpu_launches']))
" | tee -a $O/r02_multi2.log
done
