set -x
python tools/microbench.py > gpurun_out/microbench_r01.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bihome_kernel|warp_fwd_plane|warp_bwd_plane|dltn_fwd|dltn_bwd|pairgen_apply|warp_fwd_nhwc|warp_bwd_generic" -c 24 -f -o gpurun_out/prof_kernels_r01 python tools/microbench.py --once > gpurun_out/once.log 2>&1
ls -la gpurun_out
