#!/bin/bash
# round 2, call A: validate lookup v2, fresh ncu of the two north-star kernels, baseline bench
set -x
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/a_smi.txt
python -m pytest tests -m gpu -x -q > $O/a_pytest_v1.log 2>&1; echo "rc=$?" >> $O/a_pytest_v1.log
ATDN_LOOKUP_V2=1 python -m pytest tests -m gpu -x -q > $O/a_pytest_v2.log 2>&1; echo "rc=$?" >> $O/a_pytest_v2.log
LK_BATCH=54 python tools/experiments/lookup_v2_ab.py 0 > $O/a_lk_v1.txt 2>&1
LK_BATCH=54 ATDN_LOOKUP_V2=1 python tools/experiments/lookup_v2_ab.py 1 > $O/a_lk_v2.txt 2>&1
cmp $O/lk_v1.bin $O/lk_v2.bin && echo IDENTICAL > $O/a_lk_cmp.txt || echo DIFFERENT > $O/a_lk_cmp.txt
rm -f $O/lk_v1.bin $O/lk_v2.bin
python tools/experiments/exp_corr.py > $O/a_exp_corr.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'corr_pyramid|corr_lookup' -c 2 -o $O/a_ncu_corr python tools/ncu_batch.py 27 1 > $O/a_ncu.log 2>&1
ATDN_LOOKUP_V2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'corr_lookup' -c 1 -o $O/a_ncu_lkv2 python tools/ncu_batch.py 27 1 >> $O/a_ncu.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/a_bench.json 2> $O/a_bench.err
