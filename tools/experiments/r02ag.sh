#!/bin/bash
# batch size A/B on one box (270 pairs per step: 5 x 54, 3 x 90, 2 x 135)
set -x
O=gpurun_out
for b in 54 90 135; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --batch-pairs $b > $O/ag_bench_b$b.json 2> $O/ag_bench_b$b.err
done
