#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -q -x -s -k "peaked_attention" > $O/at_peaked.log 2>&1; echo "rc=$?" >> $O/at_peaked.log
