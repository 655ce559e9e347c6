mkdir -p gpurun_out
# K7b (fused BatchNorm [+ residual] -> ReLU) bring-up: parity tests, the model-level suites it touches, bench with it on / off
timeout 400 python -m pytest tests/test_gpu_stem.py -q -s --timeout 300 > gpurun_out/pytest_r04b_stem.log 2>&1; echo "stem rc=$?"
grep -n "K7\|passed\|failed\|Error" gpurun_out/pytest_r04b_stem.log | cut -c1-300 | tail -30
timeout 900 python -m pytest tests/test_gpu_zz_config0.py tests/test_gpu_zz_train_configs.py tests/test_gpu_zz_graph_and_fold.py tests/test_gpu_eval.py tests/test_gpu_head.py tests/test_gpu_zzz_field_head.py -q --timeout 600 -x > gpurun_out/pytest_r04b_model.log 2>&1; echo "model rc=$?"
tail -12 gpurun_out/pytest_r04b_model.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r04b_bench_fused.json 2> gpurun_out/r04b_bench_fused.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r04b_bench_fused.json'))
print(d['value'], d['ms_per_step'], d['layout'], d['e2e'])
for k,v in d['kernels'].items(): print(k, round(v['avg_ms']*1e3,1),'us', v['launches'], round(v['ms_per_step'],3), round(v.get('frac_of_hbm_peak',0),3))
P
BH_BNACT=aten timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r04b_bench_aten.json 2> gpurun_out/r04b_bench_aten.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r04b_bench_aten.json')); print(d['value'], d['ms_per_step'], d['layout'])"
