#!/bin/bash
# final tree on one 8-GPU box: N = 1 (no CPU baseline: it takes 10 s of 8 GPUs), N = 8, N = 4, N = 2
set -x
O=gpurun_out
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/al_bench_n1.json 2> $O/al_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > $O/al_bench_n8.json 2> $O/al_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 6 --warmup 3 --no-extras > $O/al_bench_n4.json 2> $O/al_bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 4 --warmup 3 --no-extras > $O/al_bench_n2.json 2> $O/al_bench_n2.err
