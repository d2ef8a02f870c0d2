#!/bin/bash
set -x
O=gpurun_out
for d in 4 9 2 1; do
  ATDN_FIRST_BATCH_DIV=$d python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > $O/aj_bench_div$d.json 2> $O/aj_bench_div$d.err
done
