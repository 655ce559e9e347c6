mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_head.py tests/test_gpu_fullsize.py -q -k "warp or pooled or pairgen or head" --timeout 300 > gpurun_out/pytest_r02g.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_r02g.log
timeout 600 python tools/microbench.py --warp-only > gpurun_out/microbench_warp_r02g.jsonl 2>&1; echo "microbench rc=$?"
grep -v ring gpurun_out/microbench_warp_r02g.jsonl | cut -c1-250
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/pairgen_r02g.txt
import sys, json
sys.path.insert(0, 'tools'); sys.path.insert(0, '.')
import torch, microbench
t = microbench.Timer(20)
microbench.bench_small(256, 128, t)
PY
