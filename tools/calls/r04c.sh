mkdir -p gpurun_out
timeout 300 python tools/microbench_bn.py > gpurun_out/r04c_microbench_bn.jsonl 2> gpurun_out/r04c_mb.err; echo "mb rc=$?"
timeout 300 python tools/microbench_bn.py --warm > gpurun_out/r04c_microbench_bn_warm.jsonl 2>> gpurun_out/r04c_mb.err; echo "mb rc=$?"
python - <<'P'
import json
for f in ('gpurun_out/r04c_microbench_bn.jsonl','gpurun_out/r04c_microbench_bn_warm.jsonl'):
    print(f)
    for l in open(f):
        d=json.loads(l)
        print(d['case'], d['shape'], d.get('residual'), 'fused %.0f/%.0f aten %.0f/%.0f us'%(d['fused_fwd_us'],d['fused_bwd_us'],d['aten_fwd_us'],d['aten_bwd_us']), 'frac %.2f/%.2f'%(d.get('fwd_frac_hbm',0),d.get('bwd_frac_hbm',0)))
P
timeout 600 ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r04c_ncu_bn_launches.csv python tools/microbench_bn.py --once > /dev/null 2>> gpurun_out/r04c_mb.err; echo "ncu rc=$?"
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r04c_ncu_bn_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
for r in rows[start+2:]:
    if len(r)>mi and ('bh::' in r[ki] or 'batchnorm' in r[ki] or 'clamp' in r[ki] or 'threshold' in r[ki] or 'max_pool' in r[ki]):
        print(r[ki][:70], r[gi], round(float(r[mi].replace(',',''))/1e3,1))
P
