mkdir -p gpurun_out
for v in 11 12 13 14 15; do timeout 60 tools/probes/tma_probe $v; done 2>&1 | tee gpurun_out/tma_probe_r02d.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_head.py tests/test_gpu_fullsize.py tests/test_gpu_zz_sweep_shapes.py -q --timeout 300 > gpurun_out/pytest_r02d.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_r02d.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -k "warp_paths and tile or pooled_mask and tile" --timeout 500 > gpurun_out/memcheck_warp_r02d.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/memcheck_warp_r02d.log
timeout 600 python tools/microbench.py --warp-only > gpurun_out/microbench_warp_r02d.jsonl 2>&1; echo "microbench rc=$?"
cat gpurun_out/microbench_warp_r02d.jsonl
