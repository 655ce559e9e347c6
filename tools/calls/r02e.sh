mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "warp or pooled" --timeout 300 > gpurun_out/pytest_r02e.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_r02e.log
timeout 600 python tools/microbench.py --warp-only > gpurun_out/microbench_warp_r02e.jsonl 2>&1; echo "microbench rc=$?"
cat gpurun_out/microbench_warp_r02e.jsonl
