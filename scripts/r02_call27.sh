#!/bin/bash
O=gpurun_out
mkdir -p $O; rm -f $O/parity_report.txt
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -12 ) 2>&1 | tee $O/r02_gpu_tests.log
