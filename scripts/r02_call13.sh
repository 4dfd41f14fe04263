#!/bin/bash
O=gpurun_out
mkdir -p $O; rm -f $O/parity_report.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -12 | tee $O/r02q_pytest_gpu.log
python scripts/group_probe.py 10000000 c2 default "nosplit:SWK_SPLIT_ROWS=0" "split-sig2.5:SWK_SHARE_SIGMA=2.5" 2>&1 | tee $O/r02q_groups.log
python scripts/group_probe.py 12500000 c5 default "nosplit:SWK_SPLIT_ROWS=0" 2>&1 | tee -a $O/r02q_groups.log
python bench.py --no-cpu-baseline --no-extras --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('c2 value %.4g e2e %.4g ms %.1f launches %s' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches']))
" | tee -a $O/r02q_groups.log
