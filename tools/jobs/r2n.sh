#!/usr/bin/env bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_round2_gpu.py -q -p no:cacheprovider -k "spmm" 2>&1 | tail -2
timeout 300 python tools/count_launches.py baby > gpurun_out/r2n_launch_counts_baby.txt 2>&1; grep -v Warn gpurun_out/r2n_launch_counts_baby.txt | tail -8
timeout 300 python tools/count_launches.py tiktok > gpurun_out/r2n_launch_counts_tiktok.txt 2>&1; grep "kernels" gpurun_out/r2n_launch_counts_tiktok.txt
# ncu launch list of the bench command (after its plain run exited 0)
timeout 300 python bench.py --steps 3 --warmup 3 --quick > gpurun_out/r2n_bench_quick.json 2> gpurun_out/r2n_bench_quick.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r2n_ncu_launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/r2n_ncu.log 2>&1
wc -l gpurun_out/r2n_ncu_launches.csv
