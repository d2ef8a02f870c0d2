#!/bin/bash
set -x
O=gpurun_out
python tests/gpu_diag.py --inproc time_aggregate_mixed > $O/ay_t0.log 2>&1
ATDN_PV_DBG=1 python tests/gpu_diag.py --inproc time_aggregate_mixed > $O/ay_t1.log 2>&1
python tests/gpu_diag.py --inproc time_aggregate_mixed > $O/ay_t0b.log 2>&1
