#!/bin/bash
# round 2, GPU call 24 (1 GPU): bucketed table build in two rounds (claim, retry) -- parity, timings, launch list
set -x
O=gpurun_out/r2c24; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_partitioned.py -x -q -m gpu -k "table or bucketed or cfg4_full or cfg2_full or every_stage" > $O/pytest.log 2>&1
tail -4 $O/pytest.log
env SAGE2GPU_TABLE_BUILD=bucketed timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_bucketed.json 2> $O/cfg4_bucketed.err
env SAGE2GPU_TABLE_BUILD=bucketed timeout 900 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg2_bucketed.json 2> $O/cfg2_bucketed.err
env SAGE2GPU_TABLE_BUILD=bucketed ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"table_" -c 7 --csv --log-file $O/launches_bucketed.csv python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline --no-gather --no-cfg2 > $O/ncu_bucketed.log 2>&1
ls -la $O
