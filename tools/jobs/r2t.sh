#!/usr/bin/env bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
DIFFMM_ADAM_EAGER=1 timeout 300 python -m pytest tests/test_optim_gpu.py tests/test_epoch_gpu.py tests/test_tiktok_real_gpu.py tests/test_compat_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2t_pytest.log 2>&1
echo "pytest rc=$?"; grep -v Warning gpurun_out/r2t_pytest.log | grep -E "passed|failed|FAILED" | head -6
DIFFMM_ADAM=dmm timeout 120 python tools/tiktok_real_default_mode.py 2 2>&1 | grep -E "^seed" > gpurun_out/r2t_default_dmm.txt; cat gpurun_out/r2t_default_dmm.txt
DIFFMM_ADAM=foreach timeout 120 python tools/tiktok_real_default_mode.py 2 2>&1 | grep -E "^seed" > gpurun_out/r2t_default_foreach.txt; cat gpurun_out/r2t_default_foreach.txt
