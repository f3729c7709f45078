#!/bin/bash
# round 2, GPU call 12 (1 GPU): sanity of the tight arena (cfg5 path, sharded table, low memory), then the final bench artefacts
set -x
O=gpurun_out/r2c12; mkdir -p $O
timeout 900 python tools/run_cfg5.py --genome-bp 400000000 --steps 2 --warmup 1 --low-memory 1 --table sharded > $O/cfg5_dry_400M_n1_sharded.json 2> $O/cfg5_dry_400M_n1_sharded.err
tail -c 300 $O/cfg5_dry_400M_n1_sharded.err
bash tools/runs/r2_final1.sh
cp -r gpurun_out/r2f1 $O/
