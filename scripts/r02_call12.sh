#!/bin/bash
O=gpurun_out
mkdir -p $O; rm -f $O/parity_report.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -12 | tee $O/r02n_pytest_gpu.log
python scripts/group_probe.py 10000000 c2 default 2>&1 | tee $O/r02n_groups.log
PROBE_FLAGS=7 python scripts/group_probe.py 10000000 c2 default-outputs 2>&1 | tee -a $O/r02n_groups.log
for w in c3 c3r c4; do
  python bench.py --workload $w --no-cpu-baseline --no-extras --no-e2e --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('$w value %.4g ms %.1f' % (l['value'], l['ms_per_step']))
" | tee -a $O/r02n_groups.log
done
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:"walk_fast|unpack_rows" --csv --log-file $O/r02n_traffic_c2.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras > $O/r02n_traffic.log 2>&1
python scripts/make_traffic.py c2:fast=$O/r02n_traffic_c2.csv:10000000 2>&1 | grep -E "kernel\"|\"ms\"|dram_read|dram_write|issue|threads" | head -40
grep unpack $O/r02n_traffic_c2.csv | grep -E "dram__bytes|gpu__time" | cut -c1-220
git checkout profiles/traffic.json 2>/dev/null
