#!/bin/bash
# round 2, evidence pass on the final kernels: GPU tests, smoke, both bench arms (driver form and 2000 steps),
# phase stamps, ncu launch list of the bench, ncu --set full of the two single kernels and the svb16 decoder
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s20.json 2> gpurun_out/bench_s20.err; echo "bench20 rc=$?"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json
for f in ('bench_s20', 'bench'):
    d=json.load(open('gpurun_out/%s.json' % f)); print(f, {k: d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['executed_frac'], 'cpu', d['cpu_baseline']['value'])
print('stock', {k: v.get('value') for k, v in d['gpu_stock_baseline'].items() if isinstance(v, dict)}); print('configs', {k: (v.get('value'), v.get('impl')) for k, v in d['configs'].items()}); print('variants', {k: v.get('value') for k, v in d['variants'].items()})
print('next', {k: v.get('value') for k, v in d['next_rows'].items()}, d['next_rows']['pod5_signal_decode'].get('roofline'))
r=json.load(open('gpurun_out/bench_ref.json')); print('reference arm', r['value'], r['cpu_baseline']['cores'])"
RB200_MEGA_STAMPS=1 timeout 120 python scripts/profile_step.py 6 1024 2> gpurun_out/mega_stamps.log; tail -6 gpurun_out/mega_stamps.log | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mega_kernel -s 2 -c 1 -o gpurun_out/prof_mega -f python scripts/profile_step.py 5 1024 > gpurun_out/ncu_full.log 2>&1; echo "ncu mega rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_mega_kernel -s 2 -c 1 -o gpurun_out/prof_conv_mega -f python scripts/conv_mega_quick.py conv_s64_k9 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:svb16_decode -s 1 -c 1 -o gpurun_out/prof_svb16 -f python scripts/vbz_times.py --reads 1024 > gpurun_out/ncu_vbz.log 2>&1; echo "ncu svb16 rc=$?"
ls -la gpurun_out/*.ncu-rep
