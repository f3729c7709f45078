#!/bin/bash
# round 2, GPU call 25 (1 GPU): ncu --set full of the retry round of the bucketed table build at cfg4
set -x
O=gpurun_out/r2c25; mkdir -p $O
env SAGE2GPU_TABLE_BUILD=bucketed timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"table_insert_sorted_kernel|table_claim_kernel" -c 2 -o $O/retry_cfg4 python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline --no-gather --no-cfg2 > $O/ncu.log 2>&1
tail -3 $O/ncu.log
ncu -i $O/retry_cfg4.ncu-rep --page raw --csv > $O/retry_raw.csv 2>/dev/null
ncu -i $O/retry_cfg4.ncu-rep --page source --csv --kernel-name regex:table_insert_sorted_kernel > $O/retry_source.csv 2>/dev/null
ls -la $O
