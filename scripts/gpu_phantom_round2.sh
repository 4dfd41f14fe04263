#!/bin/bash
# Bench lines + ncu of the phantom generator with the bulk-store broadcast (default) and the plain-store one.
set -u
mkdir -p gpurun_out
for w in ph-c5 ph-c2; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_$w.json 2> gpurun_out/bench3_$w.err
  tail -c 700 gpurun_out/bench3_$w.json
done
SWK_PHANTOM_BCAST=stg python bench.py --workload ph-c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_ph-c5_stg.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/phantom_launches2.csv python scripts/phantom_sizes.py c5 > gpurun_out/ncu_launches2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'slab_broadcast|cyl_slab' -c 2 -o gpurun_out/phantom_full2 python scripts/phantom_sizes.py c5 > gpurun_out/ncu_full2.log 2>&1
ncu -i gpurun_out/phantom_full2.ncu-rep --page raw --csv > gpurun_out/phantom_full2_raw.csv 2>/dev/null
python -m pytest tests/test_phantom_gpu.py tests/test_cli_gpu.py -x -q 2>&1 | tail -3
