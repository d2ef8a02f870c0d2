#!/bin/bash
# mixed fp16 / e4m3 attention probabilities as the default: full GPU suite, timing, bench A/B on one box
set -x
O=gpurun_out
python tests/gpu_diag.py --inproc time_aggregate_mixed > $O/ad_time.log 2>&1
python -m pytest tests -m gpu -q -x > $O/ad_pytest.log 2>&1; echo "rc=$?" >> $O/ad_pytest.log
cp $O/parity_sequence.json $O/ad_parity_sequence.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/ad_bench_mixed.json 2> $O/ad_bench_mixed.err
ATDN_P_MIXED=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/ad_bench_fp16.json 2> $O/ad_bench_fp16.err
