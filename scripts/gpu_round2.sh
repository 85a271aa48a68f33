#!/bin/bash
# One gpurun call for the round-2 evidence: bench line (default flags), ncu launch list of the same command, full ncu
# captures of the convexify kernel (one full-batch launch) and of the persistent SQP kernel (148 trajectories).
mkdir -p gpurun_out
( timeout 900 python bench.py ) > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity ) > gpurun_out/r2_ncu_bench.log 2>&1
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:eval_convexify -c 1 -o gpurun_out/r2_eval_full -f python scripts/ncu_eval.py 1024 cfg2 ) > gpurun_out/r2_ncu_eval.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o gpurun_out/r2_solve_full -f python scripts/ncu_solve.py 148 ) > gpurun_out/r2_ncu_solve.log 2>&1
cat gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err; tail -2 gpurun_out/r2_ncu_eval.log gpurun_out/r2_ncu_solve.log
