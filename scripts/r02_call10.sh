#!/bin/bash
O=gpurun_out
mkdir -p $O; rm -f $O/parity_report.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -12 | tee $O/r02l_pytest_gpu.log
python scripts/group_probe.py 10000000 c2 default 2>&1 | tee $O/r02l_groups.log
for w in c3 c3r c4 c1; do
  python bench.py --workload $w --no-cpu-baseline --no-extras --no-e2e --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('$w value %.4g ms %.1f' % (l['value'], l['ms_per_step']))
" | tee -a $O/r02l_groups.log
done
for w in c3 c3r; do
  SWK_NO_SHARE=1 python bench.py --workload $w --no-cpu-baseline --no-e2e --no-extras --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('$w PRIVATE value %.4g ms %.1f' % (l['value'], l['ms_per_step']))
" | tee -a $O/r02l_groups.log
done
