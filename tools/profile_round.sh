# usage: bash tools/profile_round.sh <tag>      (run on the GPU box through gpurun; outputs under gpurun_out/)
TAG=${1:-r01}
mkdir -p gpurun_out
set -x
python tools/profile_step.py > gpurun_out/step_profile_$TAG.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bihome_kernel|warp_fwd_block|warp_bwd_block|dltn_fwd|dltn_bwd|pairgen_apply|warp_fwd_nhwc|warp_bwd_generic" -c 24 -f -o gpurun_out/prof_kernels_$TAG python tools/microbench.py --once > gpurun_out/once_$TAG.log 2>&1
ls -la gpurun_out
