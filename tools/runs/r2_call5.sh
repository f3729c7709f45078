#!/bin/bash
# round 2, GPU call 5 (2 GPUs): phase C from the traversal order, probes issued ahead, partitioned build over NCCL
set -x
O=gpurun_out/r2c5; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_partitioned.py -x -q -m gpu -k "not cfg4 and not cfg3" > $O/pytest.log 2>&1
tail -6 $O/pytest.log
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "partitioned" > $O/pytest_multi.log 2>&1
tail -4 $O/pytest_multi.log
for a in 1 2 3; do
  env SAGE2GPU_PAF_AHEAD=$a timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-gather > $O/cfg2_a$a.json 2> $O/cfg2_a$a.err
done
for a in 1 3; do
  env SAGE2GPU_PAF_AHEAD=$a SAGE2GPU_PHASE_C_TIMING=1 timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_a$a.json 2> $O/cfg4_a$a.err
done
run() { # name, nproc, args...
  n=$1; np=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np "$@" > $O/$n.json 2> $O/$n.err
}
run cfg4_n2 2 --workload cfg4 --steps 5 --warmup 3 --no-gather
run cfg2_n2 2 --workload cfg2 --steps 20 --warmup 3 --no-gather
ls -la $O
