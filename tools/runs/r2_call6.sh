#!/bin/bash
# round 2, GPU call 6 (2 GPUs): partitioned build without allocation churn, device-generated slices, small cfg5-shaped dry run
set -x
O=gpurun_out/r2c6; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_partitioned.py tests/test_gpu_parity.py -x -q -m gpu -k "partitioned or generated or every_stage or superstring" > $O/pytest.log 2>&1
tail -6 $O/pytest.log
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-gather > $O/cfg2_n1.json 2> $O/cfg2_n1.err
run() { # name, nproc, args...
  n=$1; np=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np "$@" > $O/$n.json 2> $O/$n.err
}
run cfg4_n2 2 --workload cfg4 --steps 5 --warmup 3 --no-gather --no-alt-table
run cfg2_n2 2 --workload cfg2 --steps 20 --warmup 3 --no-gather --no-alt-table
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/run_cfg5.py --genome-bp 400000000 --steps 2 --warmup 1 > $O/cfg5_dry_400M_n2.json 2> $O/cfg5_dry_400M_n2.err
tail -c 1500 $O/cfg5_dry_400M_n2.err
ls -la $O
