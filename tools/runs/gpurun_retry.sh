#!/bin/bash
# retry a gpurun call while the pod answers "no box / slot free" (exit code 3, nothing charged)
# usage: gpurun_retry.sh [--gpus N] <timeout_s> <script>
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; S=$2
for i in $(seq 1 40); do
  gpurun $GP --timeout $T -- bash $S; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry $i] no slot, sleeping 90 s"; sleep 90
done
exit 3
