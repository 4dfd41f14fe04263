#!/bin/bash
# One GPU-box visit for the phantom generator: CLI tests, bench lines, ncu launch list + full capture of its kernels.
# Usage: gpurun --timeout 1500 -- bash scripts/gpu_phantom_round.sh
set -u
mkdir -p gpurun_out
python -m pytest tests/test_cli_gpu.py tests/test_phantom_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/phantom_tests.log
cat gpurun_out/phantom_tests.log
for w in ph-c5 ph-c2 ph-c3 ph-s256; do
  python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  tail -c 1500 gpurun_out/bench_$w.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/phantom_launches.csv python scripts/phantom_sizes.py c5 s256 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'slab_broadcast|cyl_slab|sphere_fill' -c 4 -o gpurun_out/phantom_full python scripts/phantom_sizes.py c5 s256 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/phantom_full.ncu-rep --page raw --csv > gpurun_out/phantom_full_raw.csv 2>/dev/null
ls -la gpurun_out
