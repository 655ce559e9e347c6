mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_head.py -q -k "warp or pooled or pairgen or head" --timeout 300 > gpurun_out/pytest_r02f.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_r02f.log
timeout 600 python tools/microbench.py --warp-only > gpurun_out/microbench_warp_r02f.jsonl 2>&1; echo "microbench rc=$?"
grep -v ring gpurun_out/microbench_warp_r02f.jsonl
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/pairgen_r02f.txt
import sys, json
sys.path.insert(0, 'tools'); sys.path.insert(0, '.')
import torch, microbench
t = microbench.Timer(20)
microbench.bench_small(256, 128, t)
PY
