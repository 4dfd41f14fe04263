#!/bin/bash
# Bench lines of the walk on device-generated phantoms + ncu of the sphere kernel.  gpurun --timeout 1800 -- bash scripts/gpu_bench_round.sh
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 1200 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
python bench.py --workload c3r --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3r.json 2> gpurun_out/bench_c3r.err; tail -c 600 gpurun_out/bench_c3r.json; tail -3 gpurun_out/bench_c3r.err
python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 1500 gpurun_out/bench_c5.json; tail -3 gpurun_out/bench_c5.err
ncu --set full --clock-control none --import-source on -k regex:'sphere_fill' -c 1 -o gpurun_out/sphere_full python scripts/phantom_sizes.py s256 > gpurun_out/ncu_sphere.log 2>&1
ncu -i gpurun_out/sphere_full.ncu-rep --page raw --csv > gpurun_out/sphere_full_raw.csv 2>/dev/null
for w in ph-c5 ph-c3; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_$w.json 2> gpurun_out/bench2_$w.err; done
ls -la gpurun_out | head -40
