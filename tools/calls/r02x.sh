mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8_r02x.json 2> gpurun_out/bench_n8_r02x.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_n8_r02x.json'))
print('N=8', b['value'], b['ms_per_step'], b['e2e']['value'], b['layout'])
PY
grep -i "error" gpurun_out/bench_n8_r02x.err | head -5
