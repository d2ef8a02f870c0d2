#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/j_pytest.log 2>&1; echo "rc=$?" >> $O/j_pytest.log
python tools/experiments/lookup_bench.py > $O/j_lookup.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'corr_lookup' -c 1 -o $O/j_ncu_lk python tools/ncu_batch.py 27 1 > $O/j_ncu.log 2>&1
