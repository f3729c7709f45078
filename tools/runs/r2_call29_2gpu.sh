#!/bin/bash
# round 2, GPU call 29 (2 GPUs): the N=2 bench line on the final build (bucketed table shards at cfg4)
set -x
O=gpurun_out/r2c29; mkdir -p $O
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --no-cfg2 ) > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 400 $O/bench_n2.err
ls -la $O
