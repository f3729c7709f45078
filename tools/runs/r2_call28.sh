#!/bin/bash
# round 2, GPU call 28 (1 GPU): launch list of the final bench command (default workload cfg4)
set -x
O=gpurun_out/r2c28; mkdir -p $O
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/bench_plain.json 2> $O/bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/ncu.log 2>&1
ls -la $O
