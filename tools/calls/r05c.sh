mkdir -p gpurun_out
timeout 215 python -m pytest tests -q -m gpu > gpurun_out/r05c_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/r05c_pytest_gpu.log | tail -3 | cut -c1-300
grep -E "^FAILED|^ERROR" gpurun_out/r05c_pytest_gpu.log | head -10 | cut -c1-300
timeout 150 python bench.py --steps 20 --warmup 5 > gpurun_out/r05c_bench_n1.json 2> gpurun_out/r05c_bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r05c_bench_n1.json'))
print('N=1 value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'launches', d['gpu_launches'], d['clocks'])
print(' roofline', d['roofline']['entry_point'], round(d['roofline']['frac'],3), 'warp+loss', round(d['warp_loss_roofline']['frac'],3), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],1))
for k,v in d['kernels'].items(): print(k, round(v['avg_ms']*1e3,1),'us', v['launches'], round(v['ms_per_step'],3), round(v.get('frac_of_hbm_peak',0),3))
P
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r05c_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r05c_smoke.log | cut -c1-200
