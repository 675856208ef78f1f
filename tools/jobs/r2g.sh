#!/usr/bin/env bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_epoch_gpu.py tests/test_modules_gpu.py tests/test_compat_gpu.py -q -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
grep -n "FAILED\|passed\|failed\|^E  " gpurun_out/r2g_pytest.log | head -40
timeout 300 python tools/epoch_times.py baby 3 > gpurun_out/r2g_epoch_times.log 2>&1; tail -7 gpurun_out/r2g_epoch_times.log
timeout 300 python tools/epoch_times.py tiktok 3 > gpurun_out/r2g_epoch_times_tiktok.log 2>&1; tail -7 gpurun_out/r2g_epoch_times_tiktok.log
