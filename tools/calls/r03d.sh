mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_triplet_kernel.py tests/test_gpu_zz_head_mirrors.py -q --timeout 200 > gpurun_out/pytest_r03d.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_r03d.log | cut -c1-400
grep -n "B=\|C=\|h=\|w=\|Hs=\|Ws=\|Ho=\|Wo=\|nhwc=\|case=\|AssertionError: " gpurun_out/pytest_r03d.log | head -30
