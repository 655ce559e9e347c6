mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r04l_bench_n4.json 2> gpurun_out/r04l_bench_n4.err; echo "bench n4 rc=$?"
tail -2 gpurun_out/r04l_bench_n4.err | cut -c1-200
python - <<'P'
import json
d=json.load(open('gpurun_out/r04l_bench_n4.json'))
print('N=4 value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'graph', d['cuda_graph'], 'e2e', d['e2e'] and round(d['e2e']['value'],1), d['final_loss'])
P
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r04l_bench_n1.json 2> gpurun_out/r04l_bench_n1.err; echo "bench n1 rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r04l_bench_n1.json'))
print('N=1 value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value'],1))
print(' roofline', d['roofline']['entry_point'], round(d['roofline']['frac'],3), d['roofline']['traffic'], 'warp+loss', round(d['warp_loss_roofline']['frac'],3))
for k,v in d['kernels'].items(): print(k, round(v['avg_ms']*1e3,1),'us', v['launches'], round(v['ms_per_step'],3), round(v.get('frac_of_hbm_peak',0),3))
P
