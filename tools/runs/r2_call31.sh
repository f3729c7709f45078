#!/bin/bash
# round 2, GPU call 31 (1 GPU box): the reference arm as the driver launches it
set -x
O=gpurun_out/r2c31; mkdir -p $O
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 ) > $O/bench_reference.json 2> $O/bench_reference.err
tail -c 300 $O/bench_reference.err; cat $O/bench_reference.json | cut -c1-600
