#!/bin/bash
# round 2, GPU call 10 (1 GPU): lean host traversal, cfg5 path with the sharded table (world 1), parity
set -x
O=gpurun_out/r2c10; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_partitioned.py tests/test_gpu_sharded.py -x -q -m gpu -k "not cfg3 and not cfg4_sharded" > $O/pytest.log 2>&1
tail -4 $O/pytest.log
env SAGE2GPU_PHASE_C_TIMING=1 timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_n1.json 2> $O/cfg4_n1.err
timeout 900 python tools/run_cfg5.py --genome-bp 400000000 --steps 2 --warmup 1 --low-memory 1 --table sharded > $O/cfg5_dry_400M_n1_sharded.json 2> $O/cfg5_dry_400M_n1_sharded.err
tail -c 600 $O/cfg5_dry_400M_n1_sharded.err
ls -la $O
