mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zzz_field_head.py tests/test_gpu_zz_graph_and_fold.py tests/test_gpu_zz_sweep_shapes.py -q --timeout 600 > gpurun_out/pytest_r02o.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_r02o.log | cut -c1-400
cat gpurun_out/graphed_step.json | head -8
timeout 600 python tools/microbench.py --field-head > gpurun_out/microbench_fieldhead_r02o.jsonl 2>&1; echo "fh microbench rc=$?"
cut -c1-200 gpurun_out/microbench_fieldhead_r02o.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02o.json 2> gpurun_out/bench_r02o.err; echo "bench rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_r02o.json'))
print(b['value'], b['ms_per_step'], b['e2e']['value'], b['layout'])
for k,v in b['kernels'].items(): print(k, round(v['avg_ms'],4), v['launches'])
PY
