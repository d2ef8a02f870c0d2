#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -q -x -s -k "resize_aa or host_streamed or sequence_parity" > $O/an_resize.log 2>&1; echo "rc=$?" >> $O/an_resize.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/an_bench.json 2> $O/an_bench.err
