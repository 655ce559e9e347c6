mkdir -p gpurun_out
free -g | head -2; nproc
timeout 900 python tools/ref_step_cuda.py --steps 20 --warmup 5 --out gpurun_out/ref_step_cuda_r02l.json > gpurun_out/ref_step_cuda_r02l.log 2>&1; echo "ref_step_cuda rc=$?"
tail -30 gpurun_out/ref_step_cuda_r02l.log
timeout 1500 python tools/mace10k.py --pairs 10000 --batch 250 --cpu-ref 500 --out gpurun_out/mace10k_r02l.json > gpurun_out/mace10k_r02l.log 2>&1; echo "mace10k rc=$?"
tail -45 gpurun_out/mace10k_r02l.log
( time timeout 900 python bench.py --impl reference --steps 4 --warmup 1 ) > gpurun_out/bench_ref_r02l.json 2> gpurun_out/bench_ref_r02l.err; echo "bench ref rc=$?"
cut -c1-1200 gpurun_out/bench_ref_r02l.json; tail -5 gpurun_out/bench_ref_r02l.err
