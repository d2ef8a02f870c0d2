#!/bin/bash
# the A/B switches still run: a short parity subset under each
set -x
O=gpurun_out
for sw in ATDN_P_MIXED=0 ATDN_P_ROWMAJOR=1 ATDN_PV_PAIR=1 ATDN_PDL=1 ATDN_MASK32=1 ATDN_NO_FORWARD_GRAPH=1 ATDN_CORR_NO_PAIR=1; do
  env $sw timeout 300 python -m pytest tests -q -x -m gpu -k "gma_full or aggregate_full or aggregate_small_odd or host_streamed" > $O/ax_$sw.log 2>&1; echo "rc=$?" >> $O/ax_$sw.log
done
