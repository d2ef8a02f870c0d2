#!/bin/bash
# final validation of round 2: the driver's sequence (GPU tests, smoke, default bench, reference arm)
set -x
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/as_pytest.log 2>&1; echo "rc=$?" >> $O/as_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/as_smoke.log 2>&1; echo "rc=$?" >> $O/as_smoke.log
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/as_bench_ref.json 2> $O/as_bench_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/as_bench.json 2> $O/as_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/as_ncu_launches.csv python bench.py --steps 1 --warmup 3 --no-graphs --frames 55 --no-cpu-baseline --no-extras > $O/as_ncu_bench.log 2>&1
