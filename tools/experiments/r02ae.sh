#!/bin/bash
set -x
O=gpurun_out
python tests/gpu_diag.py --inproc mixed_determinism mixed_determinism_flat aggregate_mixed_partly_hot aggregate_mixed_full > $O/ae_diag.log 2>&1
python -m pytest tests -m gpu -q -k "forward_graph or host_streamed" > $O/ae_pytest.log 2>&1
ATDN_P_MIXED=0 python -m pytest tests -m gpu -q -k "forward_graph or host_streamed" > $O/ae_pytest_fp16.log 2>&1
