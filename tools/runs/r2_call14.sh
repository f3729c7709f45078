#!/bin/bash
# round 2, GPU call 14 (1 GPU): the driver's sequence on the final build -- pytest -m gpu, smoke(), bench.py
set -x
O=gpurun_out/r2c14; mkdir -p $O
( time timeout 2400 python -m pytest tests -x -q -m gpu ) > $O/pytest_all.log 2>&1
tail -5 $O/pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time timeout 1500 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
tail -c 300 $O/bench_default.err
ls -la $O
