mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_stem.py -q --timeout 300 -x > gpurun_out/pytest_r04d_stem.log 2>&1; echo "stem rc=$?"
tail -4 gpurun_out/pytest_r04d_stem.log | cut -c1-300
timeout 300 python tools/microbench_bn.py > gpurun_out/r04d_microbench_bn.jsonl 2> gpurun_out/r04d_mb.err; echo "mb rc=$?"
timeout 300 python tools/microbench_bn.py --warm > gpurun_out/r04d_microbench_bn_warm.jsonl 2>> gpurun_out/r04d_mb.err; echo "mb rc=$?"
python - <<'P'
import json
for f in ('gpurun_out/r04d_microbench_bn.jsonl','gpurun_out/r04d_microbench_bn_warm.jsonl'):
    print(f)
    for l in open(f):
        d=json.loads(l)
        print(d['case'], d['shape'], d.get('residual'), 'fused %.0f/%.0f aten %.0f/%.0f us'%(d['fused_fwd_us'],d['fused_bwd_us'],d['aten_fwd_us'],d['aten_bwd_us']), 'frac %.2f/%.2f'%(d.get('fwd_frac_hbm',0),d.get('bwd_frac_hbm',0)))
P
timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r04d_bench.json 2> gpurun_out/r04d_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r04d_bench.json'))
print(d['value'], d['ms_per_step'], d['layout'], d['e2e'])
print(d['roofline'])
for k,v in d['kernels'].items(): print(k, round(v['avg_ms']*1e3,1),'us', v['launches'], round(v['ms_per_step'],3), round(v.get('frac_of_hbm_peak',0),3))
P
