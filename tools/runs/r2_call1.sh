#!/bin/bash
# round 2, GPU call 1: digest + minhash parity, id order vs min-hash order at cfg2 / cfg4, ncu of the ordered kernel
set -x
O=gpurun_out/r2c1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "digest or minhash or cfg2_full or baseline_configs" > $O/pytest.log 2>&1
tail -5 $O/pytest.log
for ord in id minhash; do
  timeout 600 python bench.py --workload cfg2 --read-order $ord --steps 20 --warmup 3 --no-cpu-baseline --no-gather > $O/cfg2_$ord.json 2> $O/cfg2_$ord.err
done
for ord in id minhash; do
  timeout 900 python bench.py --workload cfg4 --read-order $ord --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_$ord.json 2> $O/cfg4_$ord.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_a_kernel -s 3 -c 1 -o $O/pa_minhash_cfg2 \
  python bench.py --workload cfg2 --read-order minhash --steps 2 --warmup 3 --no-cpu-baseline --no-gather > $O/ncu.log 2>&1
ls -la $O
