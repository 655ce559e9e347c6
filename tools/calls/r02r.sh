mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_r02r.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_r02r.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r02r.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r02r.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02r.json 2> gpurun_out/bench_r02r.err; echo "bench rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_r02r.json'))
print(b['value'], b['ms_per_step'], b['e2e']['value'], b['layout'])
print(b['roofline'])
print(b['warp_loss_roofline'])
for k,v in b['kernels'].items(): print(k, round(v['avg_ms'],4), v['launches'])
PY
