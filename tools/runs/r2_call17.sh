#!/bin/bash
# round 2, call 17: host traversal of phase C on the box's CPU (no GPU work); needs a dump: cp gpurun_out/phasec_cfg4.bin tools/phasec_cfg4.bin first (SAGE2GPU_DUMP_PHASE_C, tools/dump_phase_c.py)
# (SAGE2GPU_WALK_AHEAD was the prefetch distance of a prototype: no effect on the box, removed from the code afterwards)
set -x
O=gpurun_out/r2c17; mkdir -p $O
g++ -O2 -std=c++17 -pthread -o /tmp/phase_c_bench tools/phase_c_bench.cpp sage2_b200/csrc/host_phase_c.cpp || exit 1
for a in 2147483647 2 4 8 16; do
  for i in 1 2; do echo "ahead $a"; SAGE2GPU_WALK_AHEAD=$a /tmp/phase_c_bench tools/phasec_cfg4.bin 20 | grep traversal; done
done > $O/walk.log 2>&1
cat $O/walk.log
