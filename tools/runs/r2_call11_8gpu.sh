#!/bin/bash
# round 2, second 8-GPU call: config #5 at FULL size with north_star's layout (reads replicated, table sharded, probes routed
# through peer memory), then cfg4 at 8 GPUs once more (lean traversal, packed-slice upload in the host-buffer leg)
set -x
O=gpurun_out/r2c11; mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 tools/run_cfg5.py --steps 2 --warmup 1 --low-memory 1 --table sharded > $O/cfg5_full_n8_sharded.json 2> $O/cfg5_full_n8_sharded.err
tail -c 1500 $O/cfg5_full_n8_sharded.err
nvidia-smi --query-gpu=memory.used --format=csv > $O/mem_after.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload cfg4 --steps 5 --warmup 3 --no-gather --no-alt-table > $O/cfg4_n8.json 2> $O/cfg4_n8.err
tail -c 300 $O/cfg4_n8.err
ls -la $O
