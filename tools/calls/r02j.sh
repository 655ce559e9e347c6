mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_head.py tests/test_gpu_fullsize.py tests/test_gpu_zz_sweep_shapes.py -q --timeout 300 > gpurun_out/pytest_r02j.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_r02j.log
timeout 600 python tools/microbench.py --warp-only > gpurun_out/microbench_warp_r02j.jsonl 2>&1; echo "microbench rc=$?"
grep -v ring gpurun_out/microbench_warp_r02j.jsonl | cut -c1-250
