#!/bin/bash
set -x
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"tc_gemm2_kernel" -s 1 -c 1 -o $O/bd_pv python tools/experiments/ncu_pv_mixed.py > $O/bd_ncu.log 2>&1
