#!/bin/bash
O=gpurun_out
run() { # name lib
  export SPINWALK_B200_LIB=$2
  [ -z "$2" ] && unset SPINWALK_B200_LIB
  python scripts/group_probe.py 10000000 c2 "$1" 2>&1 | tee -a $O/r02m_ab.log
  for w in c4 c3r c3; do
    python bench.py --workload $w --no-cpu-baseline --no-extras --no-e2e --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('$1 $w value %.4g ms %.1f' % (l['value'], l['ms_per_step']))
" | tee -a $O/r02m_ab.log
  done
}
run current ""
run head spinwalk_b200/_variants/head.so
run split spinwalk_b200/_variants/split.so
unset SPINWALK_B200_LIB
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_fast_parity_gpu.py -m gpu -q --maxfail=12 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -6 | tee $O/r02m_pytest_gpu.log
