#!/bin/bash
# GPU check of the convexify kernel: parity tests, CUDA-event timing, one ncu --set full capture.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( timeout 200 python scripts/ncu_eval.py 1024 cfg2 ) > gpurun_out/eval_time.log 2>&1
if [ -f trajopt_b200/csrc/libtb200_mb4.so ]; then
( TB200_LIB=$PWD/trajopt_b200/csrc/libtb200_mb4.so timeout 200 python scripts/ncu_eval.py 1024 cfg2 ) > gpurun_out/eval_time_mb4.log 2>&1
fi
tail -40 gpurun_out/pytest_gpu.log; cat gpurun_out/eval_time.log gpurun_out/eval_time_mb4.log
