mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_r02z.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cudnn-benchmark > gpurun_out/bench_under_ncu_r02z.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fieldhead_|moments_|affine_acc|nhwc_coop" -c 60 -f \
    -o gpurun_out/prof_k6_r02z python tools/microbench.py --once > gpurun_out/once_r02z.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_k6_r02z.ncu-rep --page raw --csv > gpurun_out/prof_kernels_raw_r02z.csv 2> /dev/null
ls -la gpurun_out/prof_k6_r02z.ncu-rep
if [ $(stat -c %s gpurun_out/prof_k6_r02z.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/prof_k6_r02z.ncu-rep; fi
