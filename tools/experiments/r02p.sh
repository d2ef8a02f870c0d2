#!/bin/bash
set -x
O=gpurun_out
python __graft_entry__.py smoke > $O/p_smoke.log 2>&1; echo "rc=$?" >> $O/p_smoke.log
python -m pytest tests -m gpu -q -x -k "two_rank" > $O/p_pytest2.log 2>&1; echo "rc=$?" >> $O/p_pytest2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 3 --warmup 3 > $O/p_bench_n4.json 2> $O/p_bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 3 > $O/p_bench_n2.json 2> $O/p_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $O/p_bench_ref_n2.json 2> $O/p_bench_ref_n2.err
