#!/bin/bash
# round 2, final single-GPU call: default bench line, launch list of the same command, full capture of K4 at cfg4
set -x
O=gpurun_out/r2f1; mkdir -p $O
( time timeout 1500 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
tail -c 400 $O/bench_default.err
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/b_cfg4.json 2> $O/b_cfg4.err && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches_cfg4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/ncu_launch.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:phase_a_fast_kernel -s 3 -c 1 -o $O/paf_cfg4 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/ncu_full.log 2>&1
ncu -i $O/paf_cfg4.ncu-rep --page raw --csv > $O/paf_cfg4_raw.csv 2>/dev/null
ls -la $O
