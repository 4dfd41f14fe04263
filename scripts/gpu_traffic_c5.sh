#!/bin/bash
# DRAM traffic of one C5 walk launch (2.5e7 spins of the 1.25e8 per GPU: 8 s per launch, few metrics => few replays).
mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum \
    --clock-control none -k regex:walk_fast --launch-skip 1 -c 1 --csv --log-file gpurun_out/traffic_c5.csv python bench.py --workload c5 --spins 25000000 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/traffic_c5.log 2>&1
tail -12 gpurun_out/traffic_c5.csv
tail -c 600 gpurun_out/traffic_c5.log
