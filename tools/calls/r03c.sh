mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_triplet_kernel.py -q --timeout 200 -k "fuzz or nhwc or True" > gpurun_out/pytest_r03c.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_r03c.log | cut -c1-400
