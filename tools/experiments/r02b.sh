#!/bin/bash
set -x
O=gpurun_out
python tests/gpu_e2e.py sequence > $O/b_sequence.log 2>&1
python -m pytest tests -m gpu -x -q -k "aggregate or sequence" > $O/b_pytest.log 2>&1; echo "rc=$?" >> $O/b_pytest.log
