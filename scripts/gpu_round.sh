#!/bin/bash
# One gpurun call: GPU parity tests, timing, bench, ncu launch list + full captures of both kernels.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( timeout 200 python scripts/ncu_eval.py 1024 cfg2 ) > gpurun_out/eval_time.log 2>&1
( timeout 300 python scripts/time_full.py cfg2 1024 64 ) > gpurun_out/time_full.log 2>&1
( timeout 400 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_convexify -c 1 -o gpurun_out/eval_full -f python scripts/ncu_eval.py 1024 cfg2 ) > gpurun_out/ncu_eval.log 2>&1
if [ -f trajopt_b200/csrc/libtb200_prof.so ]; then
( TB200_LIB=$PWD/trajopt_b200/csrc/libtb200_prof.so timeout 200 python scripts/eval_phases.py ) > gpurun_out/eval_phases.log 2>&1
fi
if [ "$NCU_SOLVE" = 1 ]; then
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o gpurun_out/solve_full -f python scripts/ncu_solve.py 148 ) > gpurun_out/ncu_solve.log 2>&1
fi
( timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline ) > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/eval_time.log; cat gpurun_out/time_full.log; cat gpurun_out/bench.json; cat gpurun_out/eval_phases.log
