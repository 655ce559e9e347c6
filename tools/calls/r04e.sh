mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],2), 'instr', round(d['ms_per_step_instrumented'],2), 'graph', d['cuda_graph'], 'e2e', d['e2e'] and round(d['e2e']['value'],1), d['gpu_launches'])
P
}
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r04e_bench_eager.json 2> gpurun_out/r04e_bench_eager.err; echo "bench rc=$?"
show gpurun_out/r04e_bench_eager.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --cuda-graph > gpurun_out/r04e_bench_graph.json 2> gpurun_out/r04e_bench_graph.err; echo "bench rc=$?"
show gpurun_out/r04e_bench_graph.json; tail -3 gpurun_out/r04e_bench_graph.err
