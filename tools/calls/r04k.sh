mkdir -p gpurun_out
TAG=r04k
# full-set capture of the K7 stem kernels (the first capture's -c 120 ended before them) + the big K7c-size tensors are left out on purpose
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"bn_stats|bn_finalize|stem_" -c 12 -f \
    -o gpurun_out/prof_stem_$TAG python tools/microbench_bn.py --once --only-stem > gpurun_out/once_stem_$TAG.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_stem_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_stem_raw_$TAG.csv 2> /dev/null
ls -la gpurun_out/prof_stem_$TAG.ncu-rep
if [ $(stat -c %s gpurun_out/prof_stem_$TAG.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/prof_stem_$TAG.ncu-rep; fi
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r04k_bench_n1.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value'],1))
print(' roofline', d['roofline']['entry_point'], round(d['roofline']['frac'],3), 'warp+loss', round(d['warp_loss_roofline']['frac'],3), 'cpu', d['cpu_baseline'])
for k,v in d['kernels'].items(): print(k, round(v['avg_ms']*1e3,1),'us', v['launches'], round(v['ms_per_step'],3), round(v.get('frac_of_hbm_peak',0),3))
P
