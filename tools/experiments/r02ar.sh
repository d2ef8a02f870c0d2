#!/bin/bash
set -x
O=gpurun_out
for i in 1 2 3; do
  ATDN_PDL=0 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > $O/ar_nopdl_$i.json 2> $O/ar_nopdl_$i.err
  ATDN_PDL=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > $O/ar_pdl_$i.json 2> $O/ar_pdl_$i.err
done
