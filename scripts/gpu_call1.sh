#!/bin/bash
# gpurun call: GPU parity tests of the generalised QP step + phase profile + one quick bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.log 2>&1
( timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_headline_batch_matches_oracle ) > gpurun_out/c1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c1_pytest.log
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k headline ) > gpurun_out/c1_headline.log 2>&1
echo "pytest exit $?" >> gpurun_out/c1_headline.log
( timeout 300 python scripts/time_full.py cfg2 1024 0 ) > gpurun_out/c1_time_full.log 2>&1
if [ -f trajopt_b200/csrc/libtb200_prof.so ]; then
( TB200_LIB=$PWD/trajopt_b200/csrc/libtb200_prof.so timeout 200 python scripts/prof_phases.py 148 cfg2 ) > gpurun_out/c1_prof_phases.log 2>&1
fi
( timeout 300 python scripts/time_cfg.py cfg3 64 50 ) > gpurun_out/c1_time_cfg3.log 2>&1
( timeout 300 python scripts/time_cfg.py cfg4 64 40 ) > gpurun_out/c1_time_cfg4.log 2>&1
tail -15 gpurun_out/c1_pytest.log; tail -5 gpurun_out/c1_headline.log; cat gpurun_out/c1_time_full.log gpurun_out/c1_prof_phases.log gpurun_out/c1_time_cfg3.log gpurun_out/c1_time_cfg4.log
( timeout 600 python bench.py --steps 2 --warmup 1 --cpu-repeats 1 ) > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -c 3000 gpurun_out/c1_bench.json; tail -5 gpurun_out/c1_bench.err
