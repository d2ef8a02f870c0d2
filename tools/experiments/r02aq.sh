#!/bin/bash
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/aq_pytest.log 2>&1; echo "rc=$?" >> $O/aq_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/aq_bench.json 2> $O/aq_bench.err
ATDN_PDL=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/aq_bench_nopdl.json 2> $O/aq_bench_nopdl.err
