mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zzz_field_head.py tests/test_gpu_zz_sweep_shapes.py tests/test_gpu_zz_graph_and_fold.py -q --timeout 120 > gpurun_out/pytest_r02v.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_r02v.log | cut -c1-300
timeout 200 python tools/microbench.py --feature-warp > gpurun_out/microbench_featwarp_r02v.jsonl 2>&1; echo "feature warp microbench rc=$?"
cut -c1-220 gpurun_out/microbench_featwarp_r02v.jsonl
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02v.json 2> gpurun_out/bench_r02v.err; echo "bench rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_r02v.json'))
print(b['value'], b['ms_per_step'], b['e2e']['value'], b['layout'])
print(b['roofline'])
for k,v in b['kernels'].items(): print(k, round(v['avg_ms'],4), v['launches'])
PY
