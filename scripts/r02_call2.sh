#!/bin/bash
# round 2, GPU call 2: the v10 kernel (shared random stream, predicated round) — parity first, then per-scale and C2 timings
O=gpurun_out
mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $O/r02b_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=12 2>&1 | tail -60 | tee $O/r02b_pytest_gpu.log
python scripts/scale_sweep.py --modes fast --spins 2000000 --flags 7 --dup 10 --scales 0.0125,0.0283,0.0641,0.1450,0.3282,0.6309,1.0301,3.2330,10.147,37.5 2>&1 | tee $O/r02b_sweep_shared.log
python scripts/scale_sweep.py --modes fast --spins 2000000 --flags 7 --scales 0.0125,0.3282,1.0301,37.5 2>&1 | tee $O/r02b_sweep_private.log
python scripts/zslab_probe.py 2000000 c2 2>&1 | tee $O/r02b_probe_c2.log
python scripts/zslab_probe.py 1500000 c5 2>&1 | tee $O/r02b_probe_c5.log
ls -la $O | tail -8
