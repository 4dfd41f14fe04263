#!/bin/bash
O=gpurun_out
for v in u4 u2 r64u2 r64; do SPINWALK_B200_LIB=spinwalk_b200/_variants/$v.so python scripts/group_probe.py 10000000 c2 "$v" 2>&1 | tee -a $O/r02k_variants.log; done
for v in u4 r64u2; do for w in c3 c4; do SPINWALK_B200_LIB=spinwalk_b200/_variants/$v.so python bench.py --workload $w --no-cpu-baseline --no-e2e --no-extras --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('$v $w value %.4g ms %.1f' % (l['value'], l['ms_per_step']))
" | tee -a $O/r02k_variants.log; done; done
