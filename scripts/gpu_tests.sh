#!/bin/bash
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/t_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/t_pytest.log
( timeout 300 python scripts/time_cfg.py cfg3 64 50 ) > gpurun_out/t_cfg3.log 2>&1
( timeout 300 python scripts/time_cfg.py cfg4 64 40 ) > gpurun_out/t_cfg4.log 2>&1
tail -12 gpurun_out/t_pytest.log; cat gpurun_out/t_cfg3.log gpurun_out/t_cfg4.log
