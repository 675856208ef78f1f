#!/usr/bin/env bash
# N = 2: staggered modality pipelines (per-modality all-gather + adjacency build inside the pipeline) vs the joined form
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/check_multigpu.py > gpurun_out/r2v_check_multigpu_n$N.log 2>&1; echo "check rc=$?"; tail -4 gpurun_out/r2v_check_multigpu_n$N.log
for mode in 1 0; do
  DIFFMM_STAGGER=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$mode bench.py --gpus $N --steps 20 --warmup 5 --no-others --no-prop --no-variants --no-long --no-aux --no-epoch --no-cpu-baseline > gpurun_out/r2v_bench_n${N}_stagger$mode.json 2> gpurun_out/r2v_bench_n${N}_stagger$mode.err; echo "bench rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/r2v_bench_n${N}_stagger$mode.json').read().strip().splitlines()[-1])
print('stagger', $mode, d['value'], d['ms_per_step'], d['e2e']['value'])"
done
