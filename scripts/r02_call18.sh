#!/bin/bash
# COMPAT mode on the raw z slab: bit-identity tests, then per-group timing with and without it
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_fast_parity_gpu.py -m gpu -q -x -k "zslab or compat or random_cases" 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -25 | tee $O/r02v_pytest.log
PROBE_MODE=compat python scripts/group_probe.py 2000000 c2 slab "full:SWK_NO_ZSLAB=1" 2>&1 | tee $O/r02v_compat_groups.log
