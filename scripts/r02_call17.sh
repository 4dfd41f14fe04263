#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "one_walk or shared_and_private" 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -25 | tee $O/r02u_pytest.log
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:"walk_fast|unpack_rows" --csv --log-file $O/r02u_traffic_c3.csv python bench.py --workload c3 --spins 2000000 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras > $O/r02u_traffic.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:walk_fast -c 2 -o $O/r02u_full_c3 -f python bench.py --workload c3 --spins 2000000 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras > $O/r02u_full.log 2>&1
ncu -i $O/r02u_full_c3.ncu-rep --page details --csv > $O/r02u_full_c3_details.csv 2>/dev/null
ncu -i $O/r02u_full_c3.ncu-rep --page source --csv --print-source sass --launch-skip 1 --launch-count 1 > $O/r02u_full_c3_source.csv 2>/dev/null
rm -f $O/r02u_full_c3.ncu-rep
python scripts/sass_hot.py $O/r02u_full_c3_source.csv 0.01 | tee $O/r02u_sass_hot_c3.txt | head -60
