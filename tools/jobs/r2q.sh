#!/usr/bin/env bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2q_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?"; grep -n "FAILED\|passed\|failed\|^E  " gpurun_out/r2q_pytest.log | head -20
