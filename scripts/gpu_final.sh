#!/bin/bash
# Short end-of-round check: GPU parity tests, the bench line, one ncu capture of the convexify kernel, the launch list.
mkdir -p gpurun_out
( timeout 400 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( timeout 300 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
( timeout 200 ncu --set full --clock-control none --import-source on -k regex:eval_convexify -c 1 -o gpurun_out/eval_full -f python scripts/ncu_eval.py 1024 cfg2 ) > gpurun_out/ncu_eval.log 2>&1
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline ) > gpurun_out/ncu_bench.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
