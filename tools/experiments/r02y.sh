#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/y_pytest.log 2>&1; echo "rc=$?" >> $O/y_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/y_bench.json 2> $O/y_bench.err
ATDN_P_ROWMAJOR=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/y_bench_rowmajor.json 2> $O/y_bench_rowmajor.err
