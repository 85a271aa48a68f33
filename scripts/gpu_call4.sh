#!/bin/bash
# gpurun call: evaluation-kernel changes (tests + timing + ncu), ncu source-level capture of the persistent SQP kernel
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -k "convexify or golden or headline or long_lvs or cpp" ) > gpurun_out/c4_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c4_pytest.log
( timeout 200 python scripts/ncu_eval.py 1024 cfg2 ) > gpurun_out/c4_eval_time.log 2>&1
( timeout 300 python scripts/time_full.py cfg2 1024 0 ) > gpurun_out/c4_time_full.log 2>&1
( TB200_LIB=$PWD/trajopt_b200/csrc/libtb200_prof.so timeout 200 python scripts/eval_phases.py ) > gpurun_out/c4_eval_phases.log 2>&1
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:eval_convexify -c 1 -o gpurun_out/c4_eval_full -f python scripts/ncu_eval.py 1024 cfg2 ) > gpurun_out/c4_ncu_eval.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o gpurun_out/c4_solve_full -f python scripts/ncu_solve.py 148 ) > gpurun_out/c4_ncu_solve.log 2>&1
tail -5 gpurun_out/c4_pytest.log; cat gpurun_out/c4_eval_time.log gpurun_out/c4_time_full.log gpurun_out/c4_eval_phases.log; tail -3 gpurun_out/c4_ncu_eval.log gpurun_out/c4_ncu_solve.log; ls -la gpurun_out/*.ncu-rep
