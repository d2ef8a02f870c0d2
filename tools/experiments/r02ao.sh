#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -q -x -k "two_rank" > $O/ao_pytest_2gpu.log 2>&1; echo "rc=$?" >> $O/ao_pytest_2gpu.log
