#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -q -x -k "corr or half_level or gma_full or sequence" > $O/r_pytest.log 2>&1; echo "rc=$?" >> $O/r_pytest.log
python tools/experiments/exp_corr.py > $O/r_exp_corr.txt 2>&1
