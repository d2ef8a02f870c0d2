#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/l_pytest.log 2>&1; echo "rc=$?" >> $O/l_pytest.log
python tools/experiments/exp_corr.py > $O/l_exp_corr.txt 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/l_bench.json 2> $O/l_bench.err
