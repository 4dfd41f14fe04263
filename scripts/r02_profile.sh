#!/bin/bash
# round 2 evidence: launch list of the bench command, few-metric captures of the walk launches (default z-slab path and full table), full sets
O=gpurun_out
mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/r02_launches.log 2>&1
ncu --metrics $M --clock-control none -k regex:"walk_fast|unpack_rows" --csv --log-file $O/r02_traffic_c2.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras > $O/r02_traffic.log 2>&1
SWK_NO_ZSLAB=1 ncu --metrics $M --clock-control none -k regex:"walk_fast|unpack_rows" --csv --log-file $O/r02_traffic_c2_full.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras > $O/r02_traffic_full.log 2>&1
ncu --metrics $M --clock-control none -k regex:"walk_fast" --csv --log-file $O/r02_traffic_c5.csv python bench.py --workload c5 --spins 25000000 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras > $O/r02_traffic_c5.log 2>&1
SWK_NO_ZSLAB=1 ncu --metrics $M --clock-control none -k regex:"walk_fast" --csv --log-file $O/r02_traffic_c5_full.csv python bench.py --workload c5 --spins 25000000 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras > $O/r02_traffic_c5_full.log 2>&1
# full set + source: the two variants of the default path on 2e6 spins (all 50 scales), and three single scales of the SHARED variant.
# The reports are turned into CSV pages here and deleted: gpurun brings back at most 64 MiB.
ncu --set full --import-source on --clock-control none -k regex:walk_fast -o $O/r02_full_c2 -f python bench.py --spins 2000000 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras > $O/r02_full_c2.log 2>&1
ncu -i $O/r02_full_c2.ncu-rep --page raw --csv > $O/r02_full_c2_raw.csv 2>/dev/null
ncu -i $O/r02_full_c2.ncu-rep --page details --csv > $O/r02_full_c2_details.csv 2>/dev/null
ncu -i $O/r02_full_c2.ncu-rep --page source --csv --print-source sass --launch-skip 3 --launch-count 1 > $O/r02_full_c2_source_shared.csv 2>/dev/null
ncu -i $O/r02_full_c2.ncu-rep --page source --csv --print-source sass --launch-skip 2 --launch-count 1 > $O/r02_full_c2_source_private.csv 2>/dev/null
rm -f $O/r02_full_c2.ncu-rep
ncu --set full --import-source on --clock-control none -k regex:walk_fast -o $O/r02_full_scales -f python scripts/scale_sweep.py --modes fast --spins 2000000 --reps 1 --flags 7 --dup 10 --scales 0.0125,1.0301,37.5 > $O/r02_full_scales.log 2>&1
ncu -i $O/r02_full_scales.ncu-rep --page raw --csv > $O/r02_full_scales_raw.csv 2>/dev/null
ncu -i $O/r02_full_scales.ncu-rep --page details --csv > $O/r02_full_scales_details.csv 2>/dev/null
rm -f $O/r02_full_scales.ncu-rep
python scripts/make_traffic.py c2:fast=$O/r02_traffic_c2.csv:10000000 c2:fast:full=$O/r02_traffic_c2_full.csv:10000000 c5:fast=$O/r02_traffic_c5.csv:25000000 c5:fast:full=$O/r02_traffic_c5_full.csv:25000000 > $O/r02_traffic_json.log 2>&1
cp profiles/traffic.json $O/traffic.json
tail -3 $O/r02_traffic_json.log
du -sh $O; ls -la $O | tail -24
