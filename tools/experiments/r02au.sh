#!/bin/bash
set -x
O=gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/au_bench.json 2> $O/au_bench.err
