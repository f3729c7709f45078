#!/bin/bash
# round 2, GPU call 19 (1 GPU): ncu --set full of the table build's two random passes at cfg4
set -x
O=gpurun_out/r2c19; mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"table_insert_kernel|table_fill_kernel" -c 2 -o $O/table_cfg4 python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline --no-gather --no-cfg2 > $O/ncu.log 2>&1
tail -3 $O/ncu.log
ncu -i $O/table_cfg4.ncu-rep --page raw --csv > $O/table_raw.csv 2>/dev/null
ls -la $O
