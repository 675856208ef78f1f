#!/usr/bin/env bash
# round-2 final N = 1 job: smoke, full GPU suite, default bench line, reference arm, sports / scale-out / tiktok lines, launch counts, ncu launch list, ncu of the gather
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2y_pytest.log 2>&1
echo "pytest rc=$?"; grep -n "FAILED\|passed\|failed" gpurun_out/r2y_pytest.log | head
timeout 900 python bench.py > gpurun_out/r2y_bench_n1.json 2> gpurun_out/r2y_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2y_bench_reference.json 2> gpurun_out/r2y_bench_reference.err; echo "ref rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --workload sports --no-epoch --no-cpu-baseline --no-aux > gpurun_out/r2y_bench_sports.json 2> gpurun_out/r2y_bench_sports.err
timeout 400 python bench.py --steps 4 --warmup 3 --workload scaleout --no-epoch --no-cpu-baseline --no-aux > gpurun_out/r2y_bench_scaleout.json 2> gpurun_out/r2y_bench_scaleout.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload tiktok --no-cpu-baseline --no-aux > gpurun_out/r2y_bench_tiktok.json 2> gpurun_out/r2y_bench_tiktok.err
timeout 300 python tools/count_launches.py baby > gpurun_out/r2y_launch_counts_baby.txt 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --quick > gpurun_out/r2y_bench_quick.json 2> gpurun_out/r2y_bench_quick.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r2y_ncu_launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/r2y_ncu.log 2>&1
wc -l gpurun_out/r2y_ncu_launches.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:csr_gather_act_split -s 8 -c 1 -o gpurun_out/r2y_gather python tools/bench_gather.py baby > gpurun_out/r2y_ncu_gather.log 2>&1
timeout 300 python tools/bench_gather.py baby > gpurun_out/r2y_bench_gather.txt 2>&1; timeout 300 python tools/bench_gather.py sports >> gpurun_out/r2y_bench_gather.txt 2>&1
