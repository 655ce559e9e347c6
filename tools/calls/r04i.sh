mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stem.py -q --timeout 200 > gpurun_out/pytest_r04i_stem.log 2>&1; rc=$?; echo "stem rc=$rc"
tail -6 gpurun_out/pytest_r04i_stem.log | cut -c1-400
grep -n "Error\|error\|FAILED" gpurun_out/pytest_r04i_stem.log | head -10 | cut -c1-300
if [ $rc -ne 0 ]; then export BH_BNACT2=aten; echo "K7c off for the rest of this call"; fi
timeout 400 python -m pytest tests/test_gpu_zz_config0.py tests/test_gpu_zz_train_configs.py tests/test_gpu_zz_graph_and_fold.py -q --timeout 300 -x > gpurun_out/pytest_r04i_model.log 2>&1; echo "model rc=$?"
tail -3 gpurun_out/pytest_r04i_model.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r04i_bench_n1.json 2> gpurun_out/r04i_bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r04i_bench_n1.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'instr', round(d['ms_per_step_instrumented'],2), 'graph', d['cuda_graph'], 'e2e', d['e2e'] and round(d['e2e']['value'],1), d['layout'])
print(' roofline', d['roofline']['entry_point'], round(d['roofline']['frac'],3), 'warp+loss', round(d['warp_loss_roofline']['frac'],3))
for k,v in d['kernels'].items(): print(k, round(v['avg_ms']*1e3,1),'us', v['launches'], round(v['ms_per_step'],3), round(v.get('frac_of_hbm_peak',0),3))
P
