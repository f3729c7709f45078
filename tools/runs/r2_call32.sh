#!/bin/bash
# round 2, GPU call 32 (1 GPU): radix scatter with the tile staged in shared memory -- whole GPU suite, then cfg4 / cfg2 timings
set -x
O=gpurun_out/r2c32; mkdir -p $O
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gather > $O/bench.json 2> $O/bench.err; tail -c 200 $O/bench.err
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > $O/pytest_all.log 2>&1
tail -4 $O/pytest_all.log
ls -la $O
