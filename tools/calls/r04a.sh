mkdir -p gpurun_out
# K7 (fused stem) bring-up + the K4 Jacobi change: their parity tests, then a short bench line with both stem settings
timeout 300 python -m pytest tests/test_gpu_stem.py -q -s --timeout 200 > gpurun_out/pytest_r04a_stem.log 2>&1; echo "stem rc=$?"
tail -25 gpurun_out/pytest_r04a_stem.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_kernels.py -q --timeout 200 -k "dltn or dlt" > gpurun_out/pytest_r04a_dltn.log 2>&1; echo "dltn rc=$?"
tail -5 gpurun_out/pytest_r04a_dltn.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r04a_bench_fused.json 2> gpurun_out/r04a_bench_fused.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r04a_bench_fused.json'))
print(d['value'], d['ms_per_step'], d['layout'])
for k,v in d['kernels'].items(): print(k, round(v['avg_ms']*1e3,1),'us', v['launches'], round(v.get('frac_of_hbm_peak',0),3))
P
BH_STEM=aten timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r04a_bench_aten.json 2> gpurun_out/r04a_bench_aten.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r04a_bench_aten.json')); print(d['value'], d['ms_per_step'], d['layout'])"
