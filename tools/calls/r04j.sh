mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_r04j_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_r04j_gpu.log | cut -c1-300
grep -n "FAILED\|ERROR" gpurun_out/pytest_r04j_gpu.log | head -10 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r04j.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r04j.log | cut -c1-300
