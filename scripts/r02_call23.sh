#!/bin/bash
# registers per thread / blocks per SM of the MULTI kernels (one walk for all scales): C3, C3r, 2e6 spins
O=gpurun_out
mkdir -p $O; rm -f $O/r02z_multi_variants.log
for v in "" multi2 multi3 multi5 multi6; do
  lib=""; [ -n "$v" ] && lib="SPINWALK_B200_LIB=$PWD/variants/lib_$v.so"
  for wl in c3 c3r; do
  env $lib python bench.py --workload $wl --spins 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('$wl ${v:-default(4)} value %.4g ms %.2f' % (l['value'], l['ms_per_step']))
" | tee -a $O/r02z_multi_variants.log
  done
done
