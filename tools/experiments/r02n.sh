#!/bin/bash
set -x
O=gpurun_out
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/n_bench_n1.json 2> $O/n_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > $O/n_bench_n8.json 2> $O/n_bench_n8.err
