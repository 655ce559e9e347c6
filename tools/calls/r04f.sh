mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'instr', round(d['ms_per_step_instrumented'],2), 'graph', d['cuda_graph'], 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'n', d['n_gpus'])
print(' roofline', d['roofline']['entry_point'], round(d['roofline']['frac'],3), 'warp+loss', round(d['warp_loss_roofline']['frac'],3))
P
}
timeout 600 python -m pytest tests/test_gpu_zz_graph_and_fold.py tests/test_gpu_eval.py -q --timeout 500 -x > gpurun_out/pytest_r04f.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_r04f.log | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r04f_bench_n1.json 2> gpurun_out/r04f_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r04f_bench_n1.err | cut -c1-300
show gpurun_out/r04f_bench_n1.json
