#!/bin/bash
# round 2, GPU call 16 (2 GPUs): the tests a one-GPU box skips + the N=2 bench line, on the final build
set -x
O=gpurun_out/r2c16; mkdir -p $O
( time timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py tests/test_gpu_sharded.py -x -q -m gpu ) > $O/pytest_2gpu.log 2>&1
tail -4 $O/pytest_2gpu.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 400 $O/bench_n2.err
ls -la $O
