# usage (through gpurun): bash tools/sanitize.sh <tag>
# compute-sanitizer over the kernel parity tests: memcheck on every custom kernel at small shapes,
# racecheck on the kernels that stage through shared memory (warp tiles / rings, loss rings, K6 tiles).  The large-batch
# cases are deselected: the tools slow kernels down 10-100x.  Output: gpurun_out/sanitizer_<tag>.txt
TAG=${1:-r02}
OUT=gpurun_out/sanitizer_$TAG.txt
mkdir -p gpurun_out
SEL='not 4096 and not 300 and not 600 and not 257 and not full_size and not 256-64 and not sweep and not speedup'
: > $OUT
for tool in memcheck racecheck; do
  echo "==== compute-sanitizer --tool $tool" >> $OUT
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 120 \
      python -m pytest tests/test_gpu_kernels.py tests/test_gpu_triplet_kernel.py tests/test_gpu_zzz_field_head.py tests/test_gpu_head.py \
      -q -x --timeout 1400 -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_${tool}_$TAG.log 2>&1
  echo "exit code $?" >> $OUT
  grep -E "ERROR SUMMARY|passed|failed|Race reported|Invalid|hazard" gpurun_out/sanitizer_${tool}_$TAG.log | sort | uniq -c | head -20 >> $OUT
done
cat $OUT
