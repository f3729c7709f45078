#!/bin/bash
# round 2, GPU call 22 (1 GPU): bucketed table build with one record per thread
set -x
O=gpurun_out/r2c22; mkdir -p $O
env SAGE2GPU_TABLE_BUILD=bucketed timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_bucketed.json 2> $O/cfg4_bucketed.err
env SAGE2GPU_TABLE_BUILD=bucketed ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"table_" -c 12 --csv --log-file $O/launches_bucketed.csv python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline --no-gather --no-cfg2 > $O/ncu_bucketed.log 2>&1
ls -la $O
