#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_headline_batch_matches_oracle ) > gpurun_out/c5_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c5_pytest.log
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k headline ) > gpurun_out/c5_headline.log 2>&1
( timeout 200 python scripts/ncu_eval.py 1024 cfg2 ) > gpurun_out/c5_eval_time.log 2>&1
( timeout 300 python scripts/time_full.py cfg2 1024 0 ) > gpurun_out/c5_time_full.log 2>&1
( TB200_LIB=$PWD/trajopt_b200/csrc/libtb200_prof.so timeout 200 python scripts/prof_phases.py 148 cfg2 ) > gpurun_out/c5_prof_phases.log 2>&1
( TB200_LIB=$PWD/trajopt_b200/csrc/libtb200_prof.so timeout 200 python scripts/eval_phases.py ) > gpurun_out/c5_eval_phases.log 2>&1
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:eval_convexify -c 1 -o gpurun_out/c5_eval_full -f python scripts/ncu_eval.py 1024 cfg2 ) > gpurun_out/c5_ncu_eval.log 2>&1
tail -8 gpurun_out/c5_pytest.log; tail -2 gpurun_out/c5_headline.log; cat gpurun_out/c5_eval_time.log gpurun_out/c5_time_full.log gpurun_out/c5_prof_phases.log gpurun_out/c5_eval_phases.log
