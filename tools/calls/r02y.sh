mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_r02y.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/ -x -q -m gpu --timeout 300 > gpurun_out/pytest_r02y.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_r02y.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02y.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r02y.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02y.json 2> gpurun_out/bench_r02y.err; echo "bench rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_r02y.json'))
print(b['value'], b['ms_per_step'], b['e2e'], b['layout'], b['gpu_launches'])
print(b['roofline']); print(b['warp_loss_roofline']); print(b['cpu_baseline']); print(b['clocks'])
PY
