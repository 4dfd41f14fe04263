#!/bin/bash
# One GPU-box session: tests, the bench lines, launch list, DRAM traffic of the C2 launch, full ncu set on three FoV scales.
# usage (from the repo root, on the GPU box): bash scripts/gpu_profile_round.sh <tag>
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/${TAG}_pytest_gpu.log
python bench.py > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; cat $O/${TAG}_bench_c2.json; tail -2 $O/${TAG}_bench_c2.err
python bench.py --workload c3 --no-cpu-baseline > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; cat $O/${TAG}_bench_c3.json; tail -2 $O/${TAG}_bench_c3.err
python bench.py --workload c4 --no-cpu-baseline --steps 1 --warmup 1 > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err; cat $O/${TAG}_bench_c4.json; tail -2 $O/${TAG}_bench_c4.err
python bench.py --workload c1 --no-cpu-baseline > $O/${TAG}_bench_c1.json 2> $O/${TAG}_bench_c1.err; cat $O/${TAG}_bench_c1.json; tail -2 $O/${TAG}_bench_c1.err
# launch list of the bench command (per-launch durations, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_launches.log 2>&1
# DRAM traffic + cache hit rates of ONE C2 walk launch (few metrics => few replays of the 2.4 s kernel)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum \
    --clock-control none -k regex:walk_fast --launch-skip 1 -c 1 --csv --log-file $O/${TAG}_traffic_c2.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/${TAG}_traffic.log 2>&1
cat $O/${TAG}_traffic_c2.csv | tail -12
# full set + source on three FoV scales (2e6 spins each, STATS and plain variants)
ncu --set full --import-source on --clock-control none -k regex:walk_fast -o $O/${TAG}_fast_scales -f python scripts/scale_sweep.py --modes fast --spins 2000000 --reps 1 --scales 0.0125,1.0301,37.5 > $O/${TAG}_ncu_scales.log 2>&1
tail -4 $O/${TAG}_ncu_scales.log
ls -la $O
