#!/bin/bash
# brick layout / L2 fetch granularity on the full (non z-invariant path) table, same box
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "brick or zslab or neighbours or streams" 2>&1 | tail -4 | tee $O/r02r_pytest.log
export PROBE_FLAGS=512
python scripts/group_probe.py 10000000 c2 full "brick:SWK_BRICK=1" full2 2>&1 | tee $O/r02r_brick.log
for g in 32 64 128; do
  SWK_L2_FETCH=$g python scripts/group_probe.py 10000000 c2 "full-l2f$g" "brick-l2f$g:SWK_BRICK=1" 2>&1 | tee -a $O/r02r_brick.log
done
python scripts/group_probe.py 12500000 c5 full "brick:SWK_BRICK=1" 2>&1 | tee -a $O/r02r_brick.log
SWK_L2_FETCH=32 python scripts/group_probe.py 12500000 c5 "full-l2f32" "brick-l2f32:SWK_BRICK=1" 2>&1 | tee -a $O/r02r_brick.log
