#!/usr/bin/env bash
# N-GPU job: NCCL parity check (edges / adjacency / partitioned propagation vs 1 GPU) + the bench line of N ranks
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/check_multigpu.py > gpurun_out/r2x_check_multigpu_n$N.log 2>&1; echo "check rc=$?"; tail -3 gpurun_out/r2x_check_multigpu_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2x_bench_n$N.json 2> gpurun_out/r2x_bench_n$N.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2x_bench_n$N.json').read().strip().splitlines()[-1])
print('N', $N, d['value'], d['ms_per_step'], d['e2e']['value']); print(d['breakdown_ms_per_step']); print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d.get('other_workloads',{}).items()})"
