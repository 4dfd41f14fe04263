#!/bin/bash
# round 2, GPU call 1: baseline measurements before the kernel work (z-slab per scale, COMPAT rate, random parity, ncu full on the slab variant)
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/r02a_pytest_gpu.log
timeout 600 python scripts/random_parity_gpu.py 100 2>&1 | tail -15 | tee $O/r02a_random_parity.log
python scripts/scale_sweep.py --modes fast --spins 2000000 --flags 7 --scales 0.0125,0.0283,0.0641,0.1450,0.3282,0.6309,1.0301,3.2330,10.147,37.5 2>&1 | tee $O/r02a_sweep_full.log
python scripts/scale_sweep.py --modes fast --spins 2000000 --flags 263 --scales 0.0125,0.0283,0.0641,0.1450,0.3282,0.6309,1.0301,3.2330,10.147,37.5 2>&1 | tee $O/r02a_sweep_zslab.log
python scripts/scale_sweep.py --modes compat --spins 1000000 --flags 7 --scales 0.0125,0.3282,1.0301,37.5 2>&1 | tee $O/r02a_sweep_compat.log
python scripts/zslab_probe.py 2000000 c2 2>&1 | tee $O/r02a_zslab_c2.log
ncu --set full --import-source on --clock-control none -k regex:walk_fast -o $O/r02a_zslab_scales -f python scripts/scale_sweep.py --modes fast --spins 2000000 --reps 1 --flags 263 --scales 0.0125,1.0301,37.5 > $O/r02a_ncu_scales.log 2>&1
tail -4 $O/r02a_ncu_scales.log
ls -la $O
