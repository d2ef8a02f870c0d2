#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/c_pytest.log 2>&1; echo "rc=$?" >> $O/c_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/c_bench.json 2> $O/c_bench.err
