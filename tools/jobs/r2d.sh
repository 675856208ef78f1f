#!/usr/bin/env bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench_baby.json 2> gpurun_out/r2d_bench_baby.err
echo "bench rc=$?"
tail -3 gpurun_out/r2d_bench_baby.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err
echo "ref rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2d_bench_baby.json').read().strip().splitlines()[-1])
print('value=%.4g ms=%.3f e2e=%.4g launches=%d'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches']))
for k in ('long_run','variants','propagation','cpu_baseline','epoch_sec','epoch_sec_cuda_graph'):
    print(k, json.dumps(d.get(k))[:500])
print('aux', {k:(round(v.get('frac',0),3), round(v.get('avg_launch_ms',0),4)) for k,v in d.get('aux_rooflines',{}).items() if isinstance(v,dict)})
print('roofline', round(d['roofline']['frac'],3), {k:(round(v['avg_ms'],4),round(v['tflops'])) for k,v in d['roofline']['by_shape_MxNxK'].items()})
print('breakdown',{k:v['ms_per_step'] for k,v in d['breakdown_ms_per_step'].items()})
r=json.loads(open('gpurun_out/r2d_bench_ref.json').read().strip().splitlines()[-1])
print('ref value', r['value'], r['cpu_baseline']['sample'][:120])
P
