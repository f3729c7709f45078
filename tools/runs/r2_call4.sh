#!/bin/bash
# round 2, GPU call 4 (2 GPUs): NCCL / CUDA-IPC parity tests on real hardware, partitioned build scaling at N=2
set -x
O=gpurun_out/r2c4; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $O/pytest.log 2>&1
tail -8 $O/pytest.log
run() { # name, nproc, args...
  n=$1; np=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np "$@" > $O/$n.json 2> $O/$n.err
}
run cfg4_n2 2 --workload cfg4 --steps 5 --warmup 3 --no-gather
run cfg2_n2 2 --workload cfg2 --steps 20 --warmup 3 --no-gather
run cfg2_n2_repl 2 --workload cfg2 --steps 20 --warmup 3 --no-gather --replicate-stages --no-alt-table
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_n1.json 2> $O/cfg4_n1.err
ls -la $O
