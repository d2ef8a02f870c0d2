#!/bin/bash
set -x
O=gpurun_out
ATDN_PV_PAIR=1 python -m pytest tests -m gpu -q -x -k "aggregate or gma_full" > $O/s_pytest.log 2>&1; echo "rc=$?" >> $O/s_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/s_bench.json 2> $O/s_bench.err
ATDN_PV_PAIR=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/s_bench_pvpair.json 2> $O/s_bench_pvpair.err
