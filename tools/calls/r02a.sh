# first GPU contact of round 2: whole -m gpu suite (no -x, K6 tests enabled), knife-edge diagnosis, both field heads benched
mkdir -p gpurun_out
BH_TEST_UNVERIFIED=1 timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_r02a.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_r02a.log
timeout 200 python tools/debug_triplet.py aware > gpurun_out/debug_triplet.log 2>&1; tail -30 gpurun_out/debug_triplet.log
timeout 120 python -m bihome_b200.autotune 0 > gpurun_out/fieldhead_selftest_r02a.json 2>&1; cat gpurun_out/fieldhead_selftest_r02a.json
for side in fused aten; do
  timeout 400 python bench.py --field-head $side --no-cpu-baseline > gpurun_out/bench_${side}_r02a.json 2> gpurun_out/bench_${side}_r02a.err; echo "bench $side rc=$?"
  tail -c 2500 gpurun_out/bench_${side}_r02a.json
done
