#!/bin/bash
set -x
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"tc_gemm2_kernel|attn_probs_kernel" -s 2 -c 2 -o $O/ab_pv_mixed python tools/experiments/ncu_pv_mixed.py > $O/ab_ncu.log 2>&1
python tests/gpu_diag.py --inproc time_aggregate_mixed > $O/ab_time0.log 2>&1
ATDN_PV_AHEAD=4 python tests/gpu_diag.py --inproc time_aggregate_mixed > $O/ab_time4.log 2>&1
ATDN_PV_AHEAD=8 python tests/gpu_diag.py --inproc time_aggregate_mixed > $O/ab_time8.log 2>&1
