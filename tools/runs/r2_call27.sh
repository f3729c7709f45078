#!/bin/bash
# round 2, GPU call 27 (1 GPU): the driver's sequence on the build with the bucketed table + cfg2 over 4 contexts with bucketed shards
set -x
O=gpurun_out/r2c27; mkdir -p $O
( time timeout 2400 python -m pytest tests -x -q -m gpu ) > $O/pytest_all.log 2>&1
tail -5 $O/pytest_all.log
env SAGE2GPU_TABLE_BUILD=bucketed timeout 900 python -m pytest tests/test_gpu_partitioned.py tests/test_gpu_sharded.py -x -q -m gpu -k "full_size" > $O/pytest_bucketed_big.log 2>&1
tail -3 $O/pytest_bucketed_big.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time timeout 1500 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
tail -c 300 $O/bench_default.err
ls -la $O
