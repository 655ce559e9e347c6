mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_graph_and_fold.py -x -q --timeout 600 -s > gpurun_out/pytest_r02m.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_r02m.log | cut -c1-400
