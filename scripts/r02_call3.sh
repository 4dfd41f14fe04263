#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=12 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -80 | tee $O/r02c_pytest_gpu.log
python scripts/group_probe.py 2000000 c2 2>&1 | tee $O/r02c_groups_c2.log
ls -la $O | tail -4
