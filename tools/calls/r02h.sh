mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tile_kernel|pairgen_apply" -c 10 -f -o gpurun_out/prof_r02h python tools/once_warp.py 4096 > gpurun_out/once_r02h.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/once_r02h.log
ls -la gpurun_out/*.ncu-rep
