#!/bin/bash
# round 2 evidence: launch list of the bench command, few-metric captures of the walk launches (default z-slab path and full table, C2 / C5 / C3),
# traffic.json stamped with the kernel sources, full sets of the C2 kernel variants.  Reports are turned into CSV pages here and deleted
# (gpurun brings back at most 64 MiB).
O=gpurun_out
mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
B="--steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/r02_launches.log 2>&1
ncu --metrics $M --clock-control none -k regex:"walk_fast|unpack_rows" --csv --log-file $O/r02_traffic_c2.csv python bench.py $B > $O/r02_traffic.log 2>&1
SWK_NO_ZSLAB=1 ncu --metrics $M --clock-control none -k regex:"walk_fast|unpack_rows" --csv --log-file $O/r02_traffic_c2_full.csv python bench.py $B > $O/r02_traffic_full.log 2>&1
ncu --metrics $M --clock-control none -k regex:"walk_fast" --csv --log-file $O/r02_traffic_c5.csv python bench.py --workload c5 --spins 25000000 $B > $O/r02_traffic_c5.log 2>&1
SWK_NO_ZSLAB=1 ncu --metrics $M --clock-control none -k regex:"walk_fast" --csv --log-file $O/r02_traffic_c5_full.csv python bench.py --workload c5 --spins 25000000 $B > $O/r02_traffic_c5_full.log 2>&1
ncu --metrics $M --clock-control none -k regex:"walk_fast|unpack_rows" --csv --log-file $O/r02_traffic_c3.csv python bench.py --workload c3 --spins 2000000 $B > $O/r02_traffic_c3.log 2>&1
ncu --metrics $M --clock-control none -k regex:"walk_compat" --csv --log-file $O/r02_traffic_c2_compat.csv python bench.py --mode compat --spins 2000000 $B > $O/r02_traffic_compat.log 2>&1
python scripts/make_traffic.py c2:fast=$O/r02_traffic_c2.csv:10000000 c2:fast:full=$O/r02_traffic_c2_full.csv:10000000 c5:fast=$O/r02_traffic_c5.csv:25000000 c5:fast:full=$O/r02_traffic_c5_full.csv:25000000 c3:fast:full=$O/r02_traffic_c3.csv:2000000 c2:compat=$O/r02_traffic_c2_compat.csv:2000000 > $O/r02_traffic_json.log 2>&1
cp profiles/traffic.json $O/traffic.json
tail -3 $O/r02_traffic_json.log
if [ "$1" = "full" ]; then
ncu --set full --import-source on --clock-control none -k regex:walk_fast -o $O/r02_full_c2 -f python bench.py --spins 2000000 $B > $O/r02_full_c2.log 2>&1
ncu -i $O/r02_full_c2.ncu-rep --page details --csv > $O/r02_full_c2_details.csv 2>/dev/null
ncu -i $O/r02_full_c2.ncu-rep --page source --csv --print-source sass --launch-skip 3 --launch-count 1 > $O/r02_full_c2_source_shared.csv 2>/dev/null
ncu -i $O/r02_full_c2.ncu-rep --page source --csv --print-source sass --launch-skip 2 --launch-count 1 > $O/r02_full_c2_source_private.csv 2>/dev/null
rm -f $O/r02_full_c2.ncu-rep
python scripts/sass_hot.py $O/r02_full_c2_source_shared.csv 0.01 > $O/r02_sass_hot_shared.txt 2>&1
python scripts/sass_hot.py $O/r02_full_c2_source_private.csv 0.01 > $O/r02_sass_hot_private.txt 2>&1
rm -f $O/r02_full_c2_source_shared.csv $O/r02_full_c2_source_private.csv
fi
du -sh $O; ls -la $O | tail -30
