mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zzz_field_head.py tests/test_gpu_zz_sweep_shapes.py -q --timeout 600 > gpurun_out/pytest_r02q.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_r02q.log | cut -c1-400
timeout 600 python tools/microbench.py --field-head > gpurun_out/microbench_fieldhead_r02q.jsonl 2>&1; echo "fh microbench rc=$?"
grep -v ATen gpurun_out/microbench_fieldhead_r02q.jsonl | cut -c1-200
timeout 600 python tools/microbench.py --feature-warp > gpurun_out/microbench_featwarp_r02q.jsonl 2>&1; echo "feature warp microbench rc=$?"
cut -c1-220 gpurun_out/microbench_featwarp_r02q.jsonl
