#!/bin/bash
# GPU check of the persistent SQP kernel: parity tests, full-batch timing against the oracle, quantum sweep.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for q in 3 1 2 4 8; do
  ( TB200_QUANTUM=$q timeout 120 python scripts/time_full.py cfg2 1024 $([ $q = 3 ] && echo 64 || echo 0) ) > gpurun_out/time_full_q$q.log 2>&1
  echo "quantum $q:"; grep -E "cfg2 B=|status match|off by" gpurun_out/time_full_q$q.log
done
