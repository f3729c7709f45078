#!/bin/bash
# round 2, GPU call 7 (2 GPUs): C++ multi-device host, low-memory mode dry run of the cfg5 path, full gpu suite timing
set -x
O=gpurun_out/r2c7; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -m gpu > $O/pytest_cli.log 2>&1
tail -6 $O/pytest_cli.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/run_cfg5.py --genome-bp 400000000 --steps 2 --warmup 1 --low-memory > $O/cfg5_dry_400M_n2_lowmem.json 2> $O/cfg5_dry_400M_n2_lowmem.err
tail -c 600 $O/cfg5_dry_400M_n2_lowmem.err
( time timeout 2400 python -m pytest tests -x -q -m gpu ) > $O/pytest_all.log 2>&1
tail -8 $O/pytest_all.log
ls -la $O
