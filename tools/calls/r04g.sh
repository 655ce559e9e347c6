mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r04g_bench_n2.json 2> gpurun_out/r04g_bench_n2.err; echo "bench rc=$?"
tail -5 gpurun_out/r04g_bench_n2.err | cut -c1-300
python - <<'P'
import json
d=json.load(open('gpurun_out/r04g_bench_n2.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'graph', d['cuda_graph'], 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'n', d['n_gpus'], d['final_loss'])
P
