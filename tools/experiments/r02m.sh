#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/m_pytest.log 2>&1; echo "rc=$?" >> $O/m_pytest.log
python tools/experiments/lookup_bench.py > $O/m_lookup.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'corr_lookup|corr_pyramid' -c 2 -o $O/m_ncu python tools/ncu_batch.py 27 1 > $O/m_ncu.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/m_bench.json 2> $O/m_bench.err
