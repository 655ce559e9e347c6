mkdir -p gpurun_out
for v in 0 1 2 3 4 5 6 7 8 9 10; do timeout 60 tools/probes/tma_probe $v; done 2>&1 | tee gpurun_out/tma_probe_r02c.txt
