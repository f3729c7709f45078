#!/bin/bash
# round 2, GPU call 3: fast path v2 (two-list order), partitioned multi-GPU build on one GPU, shim, cfg4 golden parity, phase C timing
set -x
O=gpurun_out/r2c3; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_partitioned.py tests/test_gpu_shim.py -x -q -m gpu -k "not cfg4 and not cfg3" > $O/pytest.log 2>&1
tail -15 $O/pytest.log
for v in "fast4:SAGE2GPU_PAF_MINB=4" "gen:SAGE2GPU_PA_FAST=0"; do
  n=${v%%:*}; e=${v#*:}
  env $e timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-gather > $O/cfg2_$n.json 2> $O/cfg2_$n.err
done
env SAGE2GPU_PHASE_C_TIMING=1 timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_fast4.json 2> $O/cfg4_fast4.err
env SAGE2GPU_PA_FAST=0 timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_gen.json 2> $O/cfg4_gen.err
timeout 900 python bench.py --workload cfg4 --read-order minhash --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_fast4_minhash.json 2> $O/cfg4_fast4_minhash.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_a_fast_kernel -s 3 -c 1 -o $O/paf_cfg2 \
  python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --no-gather > $O/ncu.log 2>&1
ls -la $O
