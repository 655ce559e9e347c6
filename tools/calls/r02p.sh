mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_graph_and_fold.py -q --timeout 600 -s > gpurun_out/pytest_r02p.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_r02p.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_r02p.json 2> gpurun_out/bench_n2_r02p.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_n2_r02p.json'))
print('N=2', b['value'], b['ms_per_step'], b['e2e']['value'])
PY
tail -3 gpurun_out/bench_n2_r02p.err
