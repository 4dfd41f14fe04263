#!/bin/bash
# compute-sanitizer on small runs of every kernel family (FAST SHARED / PRIVATE / MULTI, COMPAT with and without the raw slab, re-binning forced)
O=gpurun_out
mkdir -p $O
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error: \|^---\|^$" | tail -14 | tee $O/r02_sanitize_$tool.log
done
