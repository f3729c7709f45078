#!/bin/bash
# round 2, 8-GPU call: cfg4 strong scaling at 8 (and 4) GPUs with the partitioned build, then config #5 at full size
set -x
O=gpurun_out/r2c9; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
free -g > $O/host_mem.txt; nproc >> $O/host_mem.txt
run() { # name, nproc, args...
  n=$1; np=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np "$@" > $O/$n.json 2> $O/$n.err
  tail -c 300 $O/$n.err
}
run cfg4_n8 8 --workload cfg4 --steps 5 --warmup 3 --no-gather
run cfg4_n4 4 --workload cfg4 --steps 5 --warmup 3 --no-gather --no-alt-table
# config #5: first a 1/4-size run (775 Mbp) to see the memory curve, then the full size
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/run_cfg5.py --genome-bp 775000000 --steps 2 --warmup 1 --low-memory 1 > $O/cfg5_quarter_n8.json 2> $O/cfg5_quarter_n8.err
tail -c 400 $O/cfg5_quarter_n8.err
# full size only if the quarter-size peak extrapolates below the 180 GB of a B200 (else half size)
G=$(python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c9/cfg5_quarter_n8.json").read().strip().splitlines()[-1])
    peak = d["device_bytes_in_use_peak_max_over_ranks"]
    print(3100000000 if 4.0 * peak < 168e9 else 1550000000)
except Exception:
    print(1550000000)
PY
)
echo "config #5 genome size for the big run: $G" | tee $O/cfg5_choice.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 tools/run_cfg5.py --genome-bp $G --steps 2 --warmup 1 --low-memory 1 > $O/cfg5_big_n8.json 2> $O/cfg5_big_n8.err
tail -c 1500 $O/cfg5_big_n8.err
ls -la $O
