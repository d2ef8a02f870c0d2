#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/o_pytest.log 2>&1; echo "rc=$?" >> $O/o_pytest.log
python tools/experiments/scan_bench.py > $O/o_scan.txt 2>&1
python tools/experiments/lookup_bench.py > $O/o_lookup.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/o_b1_launches.csv python tools/experiments/b1_forward.py > $O/o_b1.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/o_bench.json 2> $O/o_bench.err
