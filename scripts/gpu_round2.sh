#!/bin/bash
# One gpurun call for the round-2 evidence: the GPU test suite, the bench line (default flags), the ncu launch list of the
# same command, full ncu captures of the convexify kernel (one full-batch launch) and of the persistent SQP kernel
# (148 trajectories), the phase counters of the convexify kernel and of the QP step (profile builds, scripts/build_prof.sh).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu ) > gpurun_out/r02_pytest_gpu.log 2>&1
( timeout 900 python bench.py ) > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity ) > gpurun_out/r02_ncu_bench.log 2>&1
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:eval_convexify -s 2 -c 1 -o gpurun_out/r02_eval_full -f python scripts/ncu_eval.py 1024 cfg2 ) > gpurun_out/r02_ncu_eval.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o gpurun_out/r02_solve_full -f python scripts/ncu_solve.py 148 ) > gpurun_out/r02_ncu_solve.log 2>&1
if [ -f trajopt_b200/csrc/libtb200_prof.so ]; then
( TB200_LIB=$PWD/trajopt_b200/csrc/libtb200_prof.so timeout 200 python scripts/eval_phases.py ) > gpurun_out/r02_eval_phases.log 2>&1
( TB200_LIB=$PWD/trajopt_b200/csrc/libtb200_prof.so timeout 200 python scripts/prof_phases.py 148 cfg2 ) > gpurun_out/r02_prof_phases.log 2>&1
fi
( timeout 200 python scripts/time_full.py cfg2 1024 64 ) > gpurun_out/r02_time_full.log 2>&1
( timeout 300 python scripts/time_cfg.py cfg3 64 50; timeout 200 python scripts/time_cfg.py cfg4 64 40 ) > gpurun_out/r02_time_cfg34.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_bench.json; tail -3 gpurun_out/r02_bench.err; tail -n 2 gpurun_out/r02_ncu_eval.log; tail -n 2 gpurun_out/r02_ncu_solve.log; cat gpurun_out/r02_time_cfg34.log; tail -5 gpurun_out/r02_time_full.log
