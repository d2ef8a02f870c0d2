#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/q_pytest.log 2>&1; echo "rc=$?" >> $O/q_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/q_bench.json 2> $O/q_bench.err
ATDN_MASK32=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/q_bench_mask32.json 2> $O/q_bench_mask32.err
