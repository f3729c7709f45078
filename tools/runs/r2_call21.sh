#!/bin/bash
# round 2, GPU call 21 (1 GPU): launch list of the bucketed and of the direct table build at cfg4
set -x
O=gpurun_out/r2c21; mkdir -p $O
for how in bucketed direct; do
env SAGE2GPU_TABLE_BUILD=$how ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"table_|rs_|scan_" -c 600 --csv --log-file $O/launches_$how.csv python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline --no-gather --no-cfg2 > $O/ncu_$how.log 2>&1
done
ls -la $O
