#!/bin/bash
# mixed fp16 / e4m3 attention probabilities: stage-by-stage checks, timing, then the sequence parity and a bench A/B
set -x
O=gpurun_out
for c in aggregate_mixed_small_odd aggregate_mixed_single_cta aggregate_mixed_full aggregate_mixed_peaked time_aggregate_mixed; do
  timeout 300 python tests/gpu_diag.py --inproc $c > $O/aa_$c.log 2>&1; echo "rc=$?" >> $O/aa_$c.log
done
ATDN_P_MIXED=1 timeout 600 python -m pytest tests -m gpu -q -x -k "sequence_parity or forward_graph or smoke or aggregate" > $O/aa_pytest_mixed.log 2>&1; echo "rc=$?" >> $O/aa_pytest_mixed.log
cp $O/parity_sequence.json $O/aa_parity_sequence_mixed.json
ATDN_P_MIXED=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/aa_bench_mixed.json 2> $O/aa_bench_mixed.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/aa_bench_fp16.json 2> $O/aa_bench_fp16.err
ATDN_P_MIXED=1 ATDN_P_MIXED_SINGLE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/aa_bench_mixed_single.json 2> $O/aa_bench_mixed_single.err
