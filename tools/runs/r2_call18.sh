#!/bin/bash
# round 2, GPU call 18 (1 GPU): phase C traversal from device-sorted lists
set -x
O=gpurun_out/r2c18; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_partitioned.py tests/test_random_sweep.py tests/test_gpu_shim.py tests/test_gpu_cli.py -x -q -m gpu -k "not cfg3" > $O/pytest.log 2>&1
tail -4 $O/pytest.log
env SAGE2GPU_PHASE_C_TIMING=1 timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_n1.json 2> $O/cfg4_n1.err
grep "phase C host" $O/cfg4_n1.err | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pc_ -c 40 --csv --log-file $O/pc_launches.csv python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu-baseline --no-gather --no-cfg2 > $O/ncu.log 2>&1
ls -la $O
