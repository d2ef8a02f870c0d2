#!/bin/bash
# hi + lo plane as one N = 2 BN MMA per 32 columns: stage checks, timing, full suite, bench
set -x
O=gpurun_out
python tests/gpu_diag.py --inproc aggregate_mixed_partly_hot aggregate_mixed_partly_hot_bn64 aggregate_mixed_full aggregate_mixed_small_odd mixed_determinism time_aggregate_mixed > $O/ah_diag.log 2>&1
python -m pytest tests -m gpu -q -x > $O/ah_pytest.log 2>&1; echo "rc=$?" >> $O/ah_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/ah_bench.json 2> $O/ah_bench.err
