mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_r02k.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_r02k.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r02k.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r02k.log
timeout 600 python tools/microbench.py --triplet > gpurun_out/microbench_triplet_r02k.jsonl 2>&1; echo "triplet microbench rc=$?"
cat gpurun_out/microbench_triplet_r02k.jsonl | cut -c1-220
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02k.json 2> gpurun_out/bench_r02k.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/bench_r02k.json
