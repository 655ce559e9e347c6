mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cudnn-benchmark > gpurun_out/bench_plain_r03b.json 2> gpurun_out/bench_plain_r03b.err; echo "plain bench rc=$?"
cut -c1-200 gpurun_out/bench_plain_r03b.json; tail -3 gpurun_out/bench_plain_r03b.err
BH_FIELD_HEAD=fused timeout 900 ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_r03b.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cudnn-benchmark > gpurun_out/bench_under_ncu_r03b.log 2>&1; echo "launch list rc=$?"
tail -c 600 gpurun_out/bench_under_ncu_r03b.log; wc -l gpurun_out/launches_r03b.csv
