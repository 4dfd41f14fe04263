#!/bin/bash
O=gpurun_out
mkdir -p $O; rm -f $O/parity_report.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | grep -v "^xyz\|^scale2grid\|^FoV\|^spin =\|^timepoint\|^ind =\|^MatrixSize\|^PhantomSize\|^Error\|^---\|^$" | tail -40 | tee $O/r02d_pytest_gpu.log
cat $O/parity_report.txt
python scripts/group_probe.py 2000000 c2 default "sig2:SWK_SHARE_SIGMA=2" "sig0.7:SWK_SHARE_SIGMA=0.7" "private:SWK_NO_SHARE=1" 2>&1 | tee $O/r02d_groups_c2.log
for v in priv_mb6 priv_mb8; do SPINWALK_B200_LIB=spinwalk_b200/_variants/$v.so python scripts/group_probe.py 2000000 c2 "$v:SWK_NO_SHARE=1" 2>&1 | tee -a $O/r02d_groups_c2.log; done
SPINWALK_B200_LIB=spinwalk_b200/_variants/sh_mb5.so python scripts/group_probe.py 2000000 c2 "sh_mb5-all:SWK_SHARE_SIGMA=1e9" "sh_mb5-default" 2>&1 | tee -a $O/r02d_groups_c2.log
PROBE_FLAGS=512 python scripts/group_probe.py 2000000 c2 "full-default" "full-private:SWK_NO_SHARE=1" "full-shared:SWK_SHARE_SIGMA=1e9" 2>&1 | tee $O/r02d_groups_c2_full.log
python scripts/group_probe.py 1500000 c5 default "private:SWK_NO_SHARE=1" 2>&1 | tee $O/r02d_groups_c5.log
ls -la $O | tail -4
