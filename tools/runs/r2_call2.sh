#!/bin/bash
# round 2, GPU call 2: the superstring fast path of phase A -- parity, memcheck, speed
set -x
O=gpurun_out/r2c2; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "every_stage or general_kernel or superstring or minhash or digest or baseline_configs or cfg2_full" > $O/pytest.log 2>&1
tail -15 $O/pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/memcheck.log 2>&1; echo memcheck rc=$? >> $O/memcheck.log
tail -5 $O/memcheck.log
for v in "fast4:SAGE2GPU_PAF_MINB=4" "fast5:SAGE2GPU_PAF_MINB=5" "fast3:SAGE2GPU_PAF_MINB=3" "gen:SAGE2GPU_PA_FAST=0"; do
  n=${v%%:*}; e=${v#*:}
  env $e timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-gather > $O/cfg2_$n.json 2> $O/cfg2_$n.err
done
env SAGE2GPU_PAF_MINB=4 timeout 600 python bench.py --workload cfg2 --read-order minhash --steps 20 --warmup 3 --no-cpu-baseline --no-gather > $O/cfg2_fast4_minhash.json 2> $O/cfg2_fast4_minhash.err
for v in "fast4:SAGE2GPU_PAF_MINB=4" "fast5:SAGE2GPU_PAF_MINB=5"; do
  n=${v%%:*}; e=${v#*:}
  env $e timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_$n.json 2> $O/cfg4_$n.err
done
env SAGE2GPU_PAF_MINB=4 timeout 900 python bench.py --workload cfg4 --read-order minhash --steps 5 --warmup 3 --no-cpu-baseline --no-gather --no-cfg2 > $O/cfg4_fast4_minhash.json 2> $O/cfg4_fast4_minhash.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_a_fast_kernel -s 3 -c 1 -o $O/paf_cfg2 \
  python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --no-gather > $O/ncu.log 2>&1
ls -la $O
