#!/bin/bash
# mixed attention probabilities at every batch size: full GPU suite, bench A/B (incl. online_b1) on one box
set -x
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/af_pytest.log 2>&1; echo "rc=$?" >> $O/af_pytest.log
cp $O/parity_sequence.json $O/af_parity_sequence.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/af_bench_mixed.json 2> $O/af_bench_mixed.err
ATDN_P_MIXED=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/af_bench_fp16.json 2> $O/af_bench_fp16.err
