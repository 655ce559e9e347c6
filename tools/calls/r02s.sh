mkdir -p gpurun_out
for flag in "" "--cudnn-benchmark"; do
timeout 600 python bench.py --steps 20 --warmup 8 --no-cpu-baseline --no-e2e $flag > gpurun_out/bench_r02s.json 2> gpurun_out/bench_r02s.err; echo "bench [$flag] rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_r02s.json'))
print(b['value'], b['ms_per_step'], b['layout'])
PY
done
