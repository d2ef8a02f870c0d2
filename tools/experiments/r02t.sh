#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/t_pytest.log 2>&1; echo "rc=$?" >> $O/t_pytest.log
python tools/experiments/lookup_bench.py > $O/t_lookup.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'corr_lookup' -c 1 -o $O/t_ncu python tools/ncu_batch.py 27 1 > $O/t_ncu.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/t_bench.json 2> $O/t_bench.err
