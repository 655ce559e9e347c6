mkdir -p gpurun_out
timeout 100 python tools/profile_copies.py 256 > gpurun_out/r05b_copies.txt 2> gpurun_out/r05b_copies.err; echo "copies rc=$?"
head -64 gpurun_out/r05b_copies.txt | cut -c1-420
tail -3 gpurun_out/r05b_copies.err | cut -c1-300
