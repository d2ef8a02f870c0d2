#!/bin/bash
set -x
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"attn_probs_kernel" -s 1 -c 1 -o $O/ba_attn python tools/experiments/ncu_pv_mixed.py > $O/ba_ncu.log 2>&1
