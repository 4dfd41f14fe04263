#!/bin/bash
O=gpurun_out
mkdir -p $O; rm -f $O/parity_report.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -12 | tee $O/r02i_pytest_gpu.log
for w in c2 c3 c3r c4 c1; do
  python bench.py --workload $w --no-cpu-baseline --no-extras --steps 2 --warmup 1 > $O/r02i_bench_$w.json 2> $O/r02i_bench_$w.err
  python -c "
import json
l=json.load(open('$O/r02i_bench_$w.json'))
print('$w value %.4g e2e %.4g ms %.1f launches %s lost %s' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches'], l['lost_spins']))
"
  tail -2 $O/r02i_bench_$w.err
done
for w in c3 c3r; do
  SWK_NO_SHARE=1 python bench.py --workload $w --no-cpu-baseline --no-e2e --no-extras --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('$w PRIVATE value %.4g ms %.1f' % (l['value'], l['ms_per_step']))
"
done
