#!/usr/bin/env bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/bench_spmm.py > gpurun_out/r2h_spmm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:spmm64_planned -s 3 -c 1 -o gpurun_out/r2h_spmm python tools/bench_spmm.py > gpurun_out/r2h_spmm_ncu.log 2>&1
cat gpurun_out/r2h_spmm_plain.log
ls -la gpurun_out/r2h_spmm.ncu-rep
