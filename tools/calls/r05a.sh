mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_stem.py -x -q -m gpu -k "bias or residual_blocks" > gpurun_out/r05a_pytest_k8.log 2>&1; echo "pytest k8 rc=$?"
tail -4 gpurun_out/r05a_pytest_k8.log | cut -c1-300
timeout 100 python tools/profile_copies.py 256 > gpurun_out/r05a_copies.txt 2> gpurun_out/r05a_copies.err; echo "copies rc=$?"
head -50 gpurun_out/r05a_copies.txt | cut -c1-260
timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r05a_bench_n1.json 2> gpurun_out/r05a_bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r05a_bench_n1.json'))
print('N=1 value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), d['layout'])
for k,v in d['kernels'].items(): print(k, round(v['avg_ms']*1e3,1),'us', v['launches'], round(v['ms_per_step'],3), round(v.get('frac_of_hbm_peak',0),3))
P
