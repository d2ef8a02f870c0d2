#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/g_pytest.log 2>&1; echo "rc=$?" >> $O/g_pytest.log
python tools/experiments/exp_corr.py > $O/g_exp_corr.txt 2>&1
for g in 32 128; do ATDN_L2_GRAN=$g timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum --clock-control none -k regex:'corr_lookup' -c 1 --csv --log-file $O/g_lk_gran$g.csv python tools/ncu_batch.py 27 1 > $O/g_ncu_gran$g.log 2>&1; done
python bench.py --steps 5 --warmup 3 > $O/g_bench.json 2> $O/g_bench.err
