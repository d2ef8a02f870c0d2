#!/bin/bash
# programmatic dependent launch (ATDN_PDL=1): correctness under graphs and eager, bench A/B on one box
set -x
O=gpurun_out
ATDN_PDL=1 timeout 900 python -m pytest tests -m gpu -q -x > $O/ap_pytest_pdl.log 2>&1; echo "rc=$?" >> $O/ap_pytest_pdl.log
ATDN_PDL=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/ap_bench_pdl.json 2> $O/ap_bench_pdl.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/ap_bench_nopdl.json 2> $O/ap_bench_nopdl.err
