#!/bin/bash
# final tree: the default bench line (all extras, CPU baseline) and the reference arm
set -x
O=gpurun_out
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bb_bench_ref.json 2> $O/bb_bench_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bb_bench.json 2> $O/bb_bench.err
