#!/usr/bin/env bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_kernels_gpu.py tests/test_precision_fullsize_gpu.py tests/test_fullsize_gpu.py -q -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -6 gpurun_out/r2c_pytest.log
timeout 600 python tools/bench_topk2.py > gpurun_out/r2c_topk.log 2>&1
cat gpurun_out/r2c_topk.log
timeout 300 python bench.py --steps 10 --warmup 3 --sampling-step 0 --no-epoch --no-cpu-baseline --no-aux > gpurun_out/r2c_bench_baby_ss0.json 2> gpurun_out/r2c_bench_baby_ss0.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload sports --no-epoch --no-cpu-baseline --no-aux > gpurun_out/r2c_bench_sports.json 2> gpurun_out/r2c_bench_sports.err
timeout 400 python bench.py --steps 4 --warmup 3 --workload scaleout --no-epoch --no-cpu-baseline --no-aux > gpurun_out/r2c_bench_scaleout.json 2> gpurun_out/r2c_bench_scaleout.err
python - <<'P'
import json
for f in ['baby_ss0','sports','scaleout']:
    try:
        d=json.loads(open(f'gpurun_out/r2c_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, 'value=%.4g ms=%.3f'%(d['value'],d['ms_per_step']), {k:v['ms_per_step'] for k,v in d['breakdown_ms_per_step'].items()})
    except Exception as e: print(f,'ERR',e)
P
