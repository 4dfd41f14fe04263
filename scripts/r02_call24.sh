#!/bin/bash
# blocks per SM of the SHARED variant (3 / 4 / 5 = 64 / 48 / 40 registers)
O=gpurun_out
mkdir -p $O
python scripts/group_probe.py 10000000 c2 default 2>&1 | tee $O/r02z_shared_variants.log
for v in shared3 shared5; do
  SPINWALK_B200_LIB=$PWD/variants/lib_$v.so python scripts/group_probe.py 10000000 c2 $v 2>&1 | tee -a $O/r02z_shared_variants.log
done
