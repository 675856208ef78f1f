#!/usr/bin/env bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for br in 0 9728 6528 4864; do
  if [ $br = 0 ]; then unset DIFFMM_BLOCK_ROWS; else export DIFFMM_BLOCK_ROWS=$br; fi
  timeout 300 python bench.py --quick --steps 20 --warmup 5 > gpurun_out/r2t_b$br.json 2> gpurun_out/r2t_b$br.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2t_b$br.json').read().strip().splitlines()[-1])
print('block_rows', '$br', d['value'], d['ms_per_step'], d['e2e']['value'], {k: round(v['avg_ms']*1e3,1) for k,v in d['roofline']['by_shape_MxNxK'].items()})"
done
