#!/bin/bash
set -x
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 5 --warmup 3 --no-extras > $O/be_bench_n4.json 2> $O/be_bench_n4.err
