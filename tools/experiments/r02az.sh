#!/bin/bash
# attn_probs with 16 epilogue warps
set -x
O=gpurun_out
timeout 600 python tests/gpu_diag.py --inproc attn_small attn_tiny attn_one_tile_plus attn_full aggregate_mixed_partly_hot aggregate_mixed_full aggregate_mixed_small_odd mixed_determinism aggregate_full time_aggregate_mixed time_attn > $O/az_diag.log 2>&1
timeout 600 python -m pytest tests -x -q -m gpu > $O/az_pytest.log 2>&1; echo "rc=$?" >> $O/az_pytest.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > $O/az_bench.json 2> $O/az_bench.err
