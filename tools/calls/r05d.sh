mkdir -p gpurun_out
timeout 80 python tools/profile_copies.py 256 > gpurun_out/r05d_copies.txt 2> gpurun_out/r05d_copies.err; echo "copies rc=$?"
head -24 gpurun_out/r05d_copies.txt | cut -c1-300
