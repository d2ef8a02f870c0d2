#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/h_pytest.log 2>&1; echo "rc=$?" >> $O/h_pytest.log
python tools/experiments/exp_corr.py > $O/h_exp_corr.txt 2>&1
python tools/experiments/lookup_bench.py > $O/h_lookup.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'corr_pyramid|corr_lookup' -c 2 -o $O/h_ncu_corr python tools/ncu_batch.py 27 1 > $O/h_ncu.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/h_bench.json 2> $O/h_bench.err
