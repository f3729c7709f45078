#!/bin/bash
# round 2, GPU call 30 (1 GPU): sanity of the last build -- table / low-memory / stage tests, smoke, one short bench
set -x
O=gpurun_out/r2c30; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu -k "table or every_stage or low_memory or cfg4_full or sharded" > $O/pytest.log 2>&1
tail -3 $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log | cut -c1-120
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/bench.json 2> $O/bench.err; tail -c 200 $O/bench.err
ls -la $O
