#!/usr/bin/env bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_round2_gpu.py -q -p no:cacheprovider > gpurun_out/r2p_pytest.log 2>&1
grep -n "FAILED\|passed\|failed\|^E  " gpurun_out/r2p_pytest.log | head -20
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench_baby.json 2> gpurun_out/r2p_bench_baby.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2p_bench_baby.json').read().strip().splitlines()[-1])
print('value=%.4g ms=%.3f e2e=%.4g'%(d['value'],d['ms_per_step'],d['e2e']['value']))
for k in ('long_run','variants','cpu_baseline','epoch_sec','epoch_sec_eager'): print(k, json.dumps(d.get(k))[:500])
print('aux', {k:(round(v.get('frac',0),3), round(v.get('avg_launch_ms',0),4)) for k,v in d.get('aux_rooflines',{}).items() if isinstance(v,dict)})
print('roofline', round(d['roofline']['frac'],3), {k:(round(v['avg_ms'],4),round(v['tflops'])) for k,v in d['roofline']['by_shape_MxNxK'].items()})
print('breakdown',{k:v['ms_per_step'] for k,v in d['breakdown_ms_per_step'].items()})
P
DIFFMM_QSAMPLE_RNG=philox timeout 200 python bench.py --steps 10 --warmup 3 --quick | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('philox full rows: ms', d['ms_per_step'])"
