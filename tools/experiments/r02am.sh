#!/bin/bash
# N = 8 with the re-measured flow time per pair (lead 47), and N = 1 on the same box
set -x
O=gpurun_out
python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $O/am_bench_n1.json 2> $O/am_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 8 --warmup 3 --no-extras > $O/am_bench_n8.json 2> $O/am_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 8 --warmup 3 --no-extras --lead-pairs 46 > $O/am_bench_n8_lead46.json 2> $O/am_bench_n8_lead46.err
