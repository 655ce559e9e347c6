mkdir -p gpurun_out
timeout 30 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 60 \
  python -m pytest tests/test_gpu_stem.py -q -x -p no:cacheprovider \
  -k "(stem_matches or bn_relu_matches or bn_bn_relu_matches) and not 64-32-32" > gpurun_out/r05f_sanitizer_memcheck_k7.log 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/r05f_sanitizer_memcheck_k7.log | sort | uniq -c | head
