#!/bin/bash
set -x
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 4 --warmup 3 > $O/bc_bench_n2.json 2> $O/bc_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $O/bc_bench_ref_n2.json 2> $O/bc_bench_ref_n2.err
