#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "corr or gma or sequence" > $O/e_pytest.log 2>&1; echo "rc=$?" >> $O/e_pytest.log
python tools/experiments/exp_corr.py > $O/e_exp_corr.txt 2>&1
for m in 0 1 2; do LK_BATCH=54 ATDN_LOOKUP_LD=$m python tools/experiments/lookup_v2_ab.py 0 >> $O/e_lk_ld.txt 2>&1; rm -f $O/lk_v1.bin; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'corr_pyramid' -c 1 -o $O/e_ncu_corr python tools/ncu_batch.py 27 1 > $O/e_ncu.log 2>&1
for m in 1 2; do ATDN_LOOKUP_LD=$m timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum --clock-control none -k regex:'corr_lookup' -c 1 --csv --log-file $O/e_lk_ld$m.csv python tools/ncu_batch.py 27 1 >> $O/e_ncu.log 2>&1; done
