mkdir -p gpurun_out
TAG=r04h
# launch list of the step (eager so that ncu sees ordinary launches; cudnn.benchmark off: its trial launches are not part of a step)
BH_FIELD_HEAD=fused timeout 900 ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --eager --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cudnn-benchmark \
    > gpurun_out/bench_under_ncu_$TAG.log 2>&1; echo "launch list rc=$?"
# full-set capture of the K7 / K7b kernels at the step's shapes
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bn_stats|bn_finalize|stem_|bnact_" -c 120 -f \
    -o gpurun_out/prof_bn_$TAG python tools/microbench_bn.py --once > gpurun_out/once_bn_$TAG.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_bn_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_bn_raw_$TAG.csv 2> /dev/null
ls -la gpurun_out/prof_bn_$TAG.ncu-rep
if [ $(stat -c %s gpurun_out/prof_bn_$TAG.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/prof_bn_$TAG.ncu-rep; fi
