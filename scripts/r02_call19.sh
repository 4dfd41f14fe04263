#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_fast_parity_gpu.py -m gpu -q -x -k "one_walk or c3 or pgse" 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -25 | tee $O/r02w_pytest.log
for wl in c3 c3r; do
  python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>$O/r02w_err.log | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('$wl value %.4g e2e %.4g ms %.2f launches %s' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches']))
" | tee -a $O/r02w_c3.log
done
for sl in 4 8 16; do
SWK_SLICES=$sl python bench.py --no-cpu-baseline --no-extras --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('c2 slices $sl value %.4g e2e %.4g ms %.1f e2e_ms %.1f' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['e2e']['ms_per_step']))
" | tee -a $O/r02w_c3.log
done
