mkdir -p gpurun_out
SEL='not 4096 and not 300 and not 600 and not 257 and not full_size and not 256-64 and not sweep and not speedup and not training_entry_point and not backbone and not 2-128-128 and not golden_reference_pipeline'
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_triplet_kernel.py tests/test_gpu_zzz_field_head.py tests/test_gpu_head.py \
    -q -x --timeout 700 -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_memcheck_r03a.log 2>&1
echo "memcheck exit code $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_r03a.log | tail -5
