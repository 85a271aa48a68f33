#!/bin/bash
# GPU check of the convexify kernel: parity tests, CUDA-event timing, one ncu --set full capture.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( timeout 200 python scripts/ncu_eval.py 1024 cfg2 ) > gpurun_out/eval_time.log 2>&1
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_convexify -c 1 -o gpurun_out/eval_full -f python scripts/ncu_eval.py 1024 cfg2 ) > gpurun_out/ncu_eval.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/eval_time.log
