#!/bin/bash
# tile configuration re-sweep of the halo kernel on the final tree (bench A/B, one box)
set -x
O=gpurun_out
run() { ATDN_HALO_CFG="$2" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/av_$1.json 2> $O/av_$1.err; }
run base ""
run c128_mt1_pair "128:1,128,1"
run c128_mt2_single "128:2,128,0"
run c192_single "192:1,192,0"
run c256_mt1_single "256:1,256,0"
run c64_mt2_pair "64:2,64,1"
run c96_mt2_single "96:2,96,0"
run base2 ""
