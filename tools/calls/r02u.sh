mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zzz_field_head.py tests/test_gpu_zz_sweep_shapes.py tests/test_gpu_zz_graph_and_fold.py -q --timeout 600 > gpurun_out/pytest_r02u.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_r02u.log | cut -c1-300
timeout 600 python tools/microbench.py --field-head > gpurun_out/microbench_fieldhead_r02u.jsonl 2>&1; echo "fh microbench rc=$?"
grep -v ATen gpurun_out/microbench_fieldhead_r02u.jsonl | grep 256 | cut -c1-200
timeout 600 python tools/microbench.py --feature-warp > gpurun_out/microbench_featwarp_r02u.jsonl 2>&1; echo "feature warp microbench rc=$?"
cut -c1-220 gpurun_out/microbench_featwarp_r02u.jsonl
python - <<'PY'
import bihome_b200.functional as F, subprocess, sys
PY
BH_COOP=1 timeout 600 python - <<'PY' > gpurun_out/microbench_featwarp_coop_r02u.jsonl 2>&1
import sys, os
sys.path.insert(0, 'tools'); sys.argv=['microbench.py']
import microbench as m
m.F.tune('warp_variant', 2)
t = m.Timer(20)
for B, P, C in ((64, 128, 64), (256, 128, 64), (64, 128, 256)):
    m.bench_feature_warp(B, P, C, t)
PY
echo "coop:"; cut -c1-220 gpurun_out/microbench_featwarp_coop_r02u.jsonl
