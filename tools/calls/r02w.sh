mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4_r02w.json 2> gpurun_out/bench_n4_r02w.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_n4_r02w.json'))
print('N=4', b['value'], b['ms_per_step'], b['e2e']['value'], b['layout'])
PY
grep -i "warn\|error" gpurun_out/bench_n4_r02w.err | head -5
