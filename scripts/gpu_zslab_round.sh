#!/bin/bash
# Evidence needed before SWK_RUN_ZSLAB may become the default for z-invariant phantoms: the bench lines with and without
# it, the launch list, and the DRAM traffic / cache hit rates / issue activity of one slab launch.
# usage (repo root, GPU box, ~6 min): bash scripts/gpu_zslab_round.sh <tag>
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/${TAG}_pytest_gpu.log
for w in c2 c1 c4; do
  extra=""; [ $w = c4 ] && extra="--steps 1 --warmup 1"
  python bench.py --workload $w --no-cpu-baseline $extra > $O/${TAG}_bench_${w}.json 2> $O/${TAG}_bench_${w}.err
  SWK_ZSLAB=1 python bench.py --workload $w --no-cpu-baseline $extra > $O/${TAG}_bench_${w}_zslab.json 2> $O/${TAG}_bench_${w}_zslab.err
  python - <<PY
import json
for t in ("", "_zslab"):
    d = json.load(open("$O/${TAG}_bench_${w}%s.json" % t)); print("$w%s" % t, d["value"], d["ms_per_step"], d["e2e"]["value"] if d.get("e2e") else None)
PY
done
SWK_ZSLAB=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c2_zslab.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_launches.log 2>&1
SWK_ZSLAB=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum \
    --clock-control none -k regex:walk_fast --launch-skip 1 -c 1 --csv --log-file $O/${TAG}_traffic_c2_zslab.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/${TAG}_traffic.log 2>&1
tail -12 $O/${TAG}_traffic_c2_zslab.csv
ls -la $O
