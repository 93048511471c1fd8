#!/bin/bash
# Runs each GPU test file in its own process (a faulting kernel must not poison the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in "$@"; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q -x --timeout=300 > gpurun_out/$n.log 2>&1
  echo "== $n exit $?"; tail -n 25 gpurun_out/$n.log
done
