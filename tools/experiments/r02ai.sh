#!/bin/bash
set -x
O=gpurun_out
python tests/gpu_diag.py --inproc aggregate_mixed_partly_hot aggregate_mixed_full aggregate_mixed_small_odd mixed_determinism time_aggregate_mixed > $O/ai_diag.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/ai_bench.json 2> $O/ai_bench.err
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "conv3x3_64to64_s1/" -c 2 -o $O/ai_conv64 python tools/ncu_batch.py 27 1 > $O/ai_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "gru_q/" -c 1 -o $O/ai_gruq python tools/ncu_batch.py 27 1 > $O/ai_ncu2.log 2>&1
