#!/bin/bash
# 2 x B200: the default bench line under torchrun (every sub-record at N = 2, north star C5 with the NCCL all-reduce of the sums)
O=gpurun_out
mkdir -p $O
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02y_bench_2gpu.json 2> $O/r02y_bench_2gpu.err ) 2>&1 | tail -4
python - <<'PY'
import json
l = json.loads(open('gpurun_out/r02y_bench_2gpu.json').read().strip().splitlines()[-1])
print('N=%d value %.4g e2e %.4g ms %.1f launches %s' % (l['n_gpus'], l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches']))
for k in ('full_table', 'compat', 'non_invariant', 'gradient_scales', 'other_configs', 'north_star'):
    v = l.get(k)
    if isinstance(v, dict) and 'value' in v: print(k, '%.4g' % v['value'], v.get('error', ''))
    else: print(k, json.dumps(v)[:300])
PY
tail -3 $O/r02y_bench_2gpu.err
