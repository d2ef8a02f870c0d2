#!/bin/bash
set -x
O=gpurun_out
for d in 0 1 2; do ATDN_PV_DBG=$d python tests/gpu_diag.py --inproc time_aggregate_mixed > $O/ac_dbg$d.log 2>&1; done
