bash tools/gpu_round.sh r02t tsbmln
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02t.json 2> gpurun_out/bench_ref_r02t.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_r02t.json
timeout 900 python tools/ref_step_cuda.py --steps 20 --warmup 5 --out gpurun_out/ref_step_cuda_r02t.json > gpurun_out/ref_step_cuda_r02t.log 2>&1; echo "ref_step_cuda rc=$?"; tail -32 gpurun_out/ref_step_cuda_r02t.log
timeout 600 python tools/microbench.py --warp-only > gpurun_out/microbench_warp_r02t.jsonl 2>&1; echo "warp microbench rc=$?"
timeout 600 python tools/microbench.py --triplet > gpurun_out/microbench_triplet_r02t.jsonl 2>&1
timeout 600 python tools/microbench.py --field-head > gpurun_out/microbench_fieldhead_r02t.jsonl 2>&1
timeout 600 python tools/microbench.py --feature-warp > gpurun_out/microbench_featwarp_r02t.jsonl 2>&1
bash tools/sanitize.sh r02t
ls -la gpurun_out | tail -40
