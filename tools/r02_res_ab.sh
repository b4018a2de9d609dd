#!/bin/bash
# Resident-row kernel (wres, default) against the scratch warp kernel (ABCOCT_KERNEL=2) and the group kernel (ABCOCT_KERNEL=1):
# GPU parity suite first, then a value + light ncu counters per configuration (tools/r02_sweep.sh).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/res_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/res_pytest.log
tail -5 gpurun_out/res_pytest.log
bash tools/r02_sweep.sh <<C
res2048 c5-2048 ABCOCT_KERNEL=0
scr2048 c5-2048 ABCOCT_KERNEL=2
res1280_16 c1 ABCOCT_KERNEL=0
res1280_12 c1 ABCOCT_WRES_NW=12
scr1280 c1 ABCOCT_KERNEL=2
resc2_16 c2 ABCOCT_KERNEL=0
resc2_12 c2 ABCOCT_WRES_NW=12
scrc2 c2 ABCOCT_KERNEL=2
resc4 c4 ABCOCT_KERNEL=0
scrc4 c4 ABCOCT_KERNEL=2
C
