# usage (through gpurun): bash tools/gpu_round.sh <tag> [stages]     outputs under gpurun_out/
# stages: any of t(ests) b(ench) r(eference arm) m(icrobench) l(aunch list) n(cu full) s(moke) u(nverified: the opt-in
# K6 tests and a bench line with the fused field head); default "tsbmln"
# Every stage runs under its own `timeout` so a hung kernel cannot hold the box.
TAG=${1:-r01}
ST=${2:-tsbmln}
mkdir -p gpurun_out
has() { case "$ST" in *$1*) return 0;; *) return 1;; esac; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi_$TAG.txt 2>&1
if has t; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
  tail -5 gpurun_out/pytest_$TAG.log
fi
if has u; then
  BH_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_zzz_field_head.py -q > gpurun_out/pytest_unverified_$TAG.log 2>&1; echo "unverified rc=$?"
  tail -5 gpurun_out/pytest_unverified_$TAG.log
  timeout 120 python -m bihome_b200.autotune 0 > gpurun_out/fieldhead_selftest_$TAG.json 2>&1; cat gpurun_out/fieldhead_selftest_$TAG.json
  timeout 300 python tools/microbench.py --field-head > gpurun_out/microbench_fieldhead_$TAG.jsonl 2>&1; cat gpurun_out/microbench_fieldhead_$TAG.jsonl
  for side in fused aten; do
    timeout 600 python bench.py --field-head $side --no-cpu-baseline > gpurun_out/bench_${side}_$TAG.json 2> gpurun_out/bench_${side}_$TAG.err; echo "bench $side rc=$?"
    tail -c 1200 gpurun_out/bench_${side}_$TAG.json
  done
fi
if has s; then
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
fi
if has b; then
  timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_$TAG.json
fi
if has r; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_$TAG.json
fi
if has m; then
  timeout 600 python tools/microbench.py > gpurun_out/microbench_$TAG.jsonl 2>&1; echo "microbench rc=$?"; head -14 gpurun_out/microbench_$TAG.jsonl
  # K7 / K7b per tensor shape of the step, against cuDNN BatchNorm + ATen ReLU / MaxPool
  timeout 300 python tools/microbench_bn.py > gpurun_out/microbench_bn_$TAG.jsonl 2>&1; echo "microbench_bn rc=$?"; tail -3 gpurun_out/microbench_bn_$TAG.jsonl
fi
if has l; then
  # application only: ncu would otherwise follow the field head's self-test into its child process; cudnn.benchmark off: its
  # trial launches are not part of a step
  BH_FIELD_HEAD=fused timeout 900 ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv \
      --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cudnn-benchmark \
      > gpurun_out/bench_under_ncu_$TAG.log 2>&1; echo "launch list rc=$?"
fi
if has n; then
  timeout 1500 ncu --set full --clock-control none --import-source on \
      -k regex:"bihome_|triplet_|warp_fwd|warp_bwd|mask_pooled|dlt4_|dltn_|pairgen_|mace_|fieldhead_|moments_|affine_acc" -c 260 -f \
      -o gpurun_out/prof_kernels_$TAG python tools/microbench.py --once > gpurun_out/once_$TAG.log 2>&1; echo "ncu full rc=$?"
  # the report itself can exceed what gpurun brings back (64 MiB): export the raw page here and keep the report only if small
  ncu -i gpurun_out/prof_kernels_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_kernels_raw_$TAG.csv 2> /dev/null
  ls -la gpurun_out/prof_kernels_$TAG.ncu-rep
  if [ $(stat -c %s gpurun_out/prof_kernels_$TAG.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/prof_kernels_$TAG.ncu-rep; fi
fi
if has n; then
  # the K7 stem kernels (268 MB tensors: ncu's replays are slow, keep the capture short); tools/calls/r04h.sh has the K7b set
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"bn_stats|bn_finalize|stem_" -c 12 -f \
      -o gpurun_out/prof_stem_$TAG python tools/microbench_bn.py --once --only-stem > gpurun_out/once_stem_$TAG.log 2>&1; echo "ncu stem rc=$?"
  ncu -i gpurun_out/prof_stem_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_stem_raw_$TAG.csv 2> /dev/null
fi
ls -la gpurun_out
