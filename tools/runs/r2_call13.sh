#!/bin/bash
# round 2, GPU call 13 (1 GPU): phase C traversal over connected components on several host threads
set -x
O=gpurun_out/r2c13; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_partitioned.py tests/test_gpu_shim.py -x -q -m gpu -k "not cfg3" > $O/pytest.log 2>&1
tail -4 $O/pytest.log
env SAGE2GPU_PHASE_C_TIMING=1 timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_n1.json 2> $O/cfg4_n1.err
grep "phase C host" $O/cfg4_n1.err | tail -2
env SAGE2GPU_PHASE_C_TIMING=1 SAGE2GPU_HOST_THREADS=4 timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_n1_t4.json 2> $O/cfg4_n1_t4.err
grep "phase C host" $O/cfg4_n1_t4.err | tail -1
ls -la $O
