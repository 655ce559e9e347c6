mkdir -p gpurun_out
timeout 52 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 60 \
  python -m pytest tests/test_gpu_stem.py -q -x -p no:cacheprovider \
  -k "(channel_bias_kernels or conv_transpose_bias_matches) and not 128-128 and not 16-128" > gpurun_out/r05e_sanitizer_memcheck_k8.log 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/r05e_sanitizer_memcheck_k8.log | sort | uniq -c | head
