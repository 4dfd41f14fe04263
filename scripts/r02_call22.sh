#!/bin/bash
# tuning variants of the same library (SPINWALK_B200_LIB): sync period of the events, registers per thread of the PRIVATE variant
O=gpurun_out
mkdir -p $O
python scripts/group_probe.py 10000000 c2 default 2>&1 | tee $O/r02z_variants.log
for v in ksync16 ksync4 minb4 minb6; do
  SPINWALK_B200_LIB=$PWD/variants/lib_$v.so python scripts/group_probe.py 10000000 c2 $v 2>&1 | tee -a $O/r02z_variants.log
done
for v in "" ksync16 ksync4; do
  lib=""; [ -n "$v" ] && lib="SPINWALK_B200_LIB=$PWD/variants/lib_$v.so"
  env $lib python bench.py --workload c4 --spins 2000000 --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('c4 $v value %.4g ms %.2f' % (l['value'], l['ms_per_step']))
" | tee -a $O/r02z_variants.log
done
