#!/usr/bin/env bash
# round-2 job B: new tests (eval, compat, round2, precision, tiktok) + top-k microbench with the per-kernel split
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/test_round2_gpu.py tests/test_compat_gpu.py tests/test_epoch_gpu.py tests/test_precision_fullsize_gpu.py tests/test_tiktok_real_gpu.py -q -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -8 gpurun_out/r2b_pytest.log
timeout 600 python tools/bench_topk2.py > gpurun_out/r2b_topk.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_topk_launches.csv python tools/bench_topk2.py baby sports > gpurun_out/r2b_topk_ncu.log 2>&1
cat gpurun_out/r2b_topk.log
timeout 300 python bench.py --steps 10 --warmup 3 --sampling-step 0 --no-epoch --no-cpu-baseline --no-aux > gpurun_out/r2b_bench_baby_ss0.json 2> gpurun_out/r2b_bench_baby_ss0.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload sports --no-epoch --no-cpu-baseline --no-aux > gpurun_out/r2b_bench_sports.json 2> gpurun_out/r2b_bench_sports.err
timeout 400 python bench.py --steps 4 --warmup 3 --workload scaleout --no-epoch --no-cpu-baseline --no-aux > gpurun_out/r2b_bench_scaleout.json 2> gpurun_out/r2b_bench_scaleout.err
