#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "two_rank or forward_graph or graphs_are_retired or corrblock" > $O/d_pytest.log 2>&1; echo "rc=$?" >> $O/d_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/d_bench_n2.json 2> $O/d_bench_n2.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/d_b1_launches.csv python tools/experiments/b1_forward.py > $O/d_b1.log 2>&1
