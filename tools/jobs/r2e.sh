#!/usr/bin/env bash
# 2 GPUs: NCCL parity of the sharded rebuild + partitioned propagation, then the N=2 bench line with every extra
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_multigpu.py > gpurun_out/r2e_check_multigpu_n2.log 2>&1
echo "check rc=$?"; tail -4 gpurun_out/r2e_check_multigpu_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2e_bench_n2.json').read().strip().splitlines()[-1])
print('value=%.4g ms=%.3f e2e=%.4g'%(d['value'],d['ms_per_step'],d['e2e']['value']))
for k in ('long_run','variants','other_workloads','propagation'):
    print(k, json.dumps(d.get(k))[:900])
print('breakdown',{k:v['ms_per_step'] for k,v in d['breakdown_ms_per_step'].items()})
P
timeout 300 python -m pytest tests/test_multigpu_gpu.py -q -p no:cacheprovider 2>&1 | tail -3
