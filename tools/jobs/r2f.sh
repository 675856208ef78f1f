#!/usr/bin/env bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_epoch_gpu.py tests/test_modules_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -25 gpurun_out/r2f_pytest.log
timeout 300 python tools/epoch_times.py baby 3 > gpurun_out/r2f_epoch_times.log 2>&1; tail -12 gpurun_out/r2f_epoch_times.log
