#!/bin/bash
set -x
O=gpurun_out
python bench.py --gpus 1 --steps 10 --warmup 3 > $O/u_bench_n1.json 2> $O/u_bench_n1.err
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/u_bench_ref.json 2> $O/u_bench_ref.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > $O/u_bench_n8.json 2> $O/u_bench_n8.err
