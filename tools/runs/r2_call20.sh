#!/bin/bash
# round 2, GPU call 20 (1 GPU): bucketed table build -- parity on both paths, then cfg2 / cfg4 timings either way
set -x
O=gpurun_out/r2c20; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_partitioned.py -x -q -m gpu -k "table or bucketed or cfg4_full or cfg2_full or every_stage" > $O/pytest.log 2>&1
tail -4 $O/pytest.log
for how in direct bucketed; do
  env SAGE2GPU_TABLE_BUILD=$how timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_$how.json 2> $O/cfg4_$how.err
  env SAGE2GPU_TABLE_BUILD=$how timeout 900 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg2_$how.json 2> $O/cfg2_$how.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"table_|rs_" -c 60 --csv --log-file $O/table_launches.csv python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline --no-gather --no-cfg2 > $O/ncu.log 2>&1
ls -la $O
