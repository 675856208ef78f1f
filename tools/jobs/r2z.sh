#!/usr/bin/env bash
# ncu --set full of one modality pipeline of a rebuild step (final round-2 code), pipelines serialised on one stream
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
DIFFMM_STREAMS=1 timeout 900 ncu --set full --import-source on --clock-control none -s 400 -c 26 -o gpurun_out/r2z_step python bench.py --quick --steps 3 --warmup 3 > gpurun_out/r2z_ncu.log 2>&1
ls -la gpurun_out/r2z_step.ncu-rep; grep -c "Profiling" gpurun_out/r2z_ncu.log
