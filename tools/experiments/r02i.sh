#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/i_pytest.log 2>&1; echo "rc=$?" >> $O/i_pytest.log
python tools/experiments/lookup_bench.py > $O/i_lookup.txt 2>&1
