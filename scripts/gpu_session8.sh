#!/bin/bash
# Validation of the round's final state on one B200 (ordered by priority; every step bounded):
# smoke, the full GPU test suite, tree and direct bench lines, launch list of one tree evaluation,
# reference arm, galaxy-model bench.
mkdir -p gpurun_out
BUDGET=${GH_SESSION_BUDGET:-690}
left() { echo $(( BUDGET - SECONDS )); }
run() {  # run <min seconds needed> <timeout> cmd...
  local need=$1 to=$2; shift 2
  if [ $(left) -lt $need ]; then echo "SKIP (only $(left) s left): $*"; return 99; fi
  [ $to -gt $(left) ] && to=$(left)
  timeout $to "$@"
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
run 60 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest (t=$SECONDS)"
run 120 420 python -m pytest tests -q -m gpu -n 3 --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log | cut -c1-220
echo "== bench tree (t=$SECONDS)"
run 60 200 python bench.py --workload tree > gpurun_out/bench_tree.json 2> gpurun_out/bench_tree.err; echo "rc=$?"; cut -c1-3000 gpurun_out/bench_tree.json; tail -3 gpurun_out/bench_tree.err
echo "== bench direct (t=$SECONDS)"
run 60 240 python bench.py > gpurun_out/bench_direct.json 2> gpurun_out/bench_direct.err; echo "rc=$?"; cut -c1-2500 gpurun_out/bench_direct.json; tail -3 gpurun_out/bench_direct.err
echo "== ncu launch list, tree (t=$SECONDS)"
run 60 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tree32.csv python scripts/profile_kernels.py tree32 4194304 > gpurun_out/ncu_tree.log 2>&1; tail -1 gpurun_out/ncu_tree.log
echo "== reference arm (t=$SECONDS)"
run 45 150 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-400 gpurun_out/bench_reference.json
echo "== bench galaxy (t=$SECONDS)"
run 90 240 python bench.py --workload galaxy --no-cpu-baseline > gpurun_out/bench_galaxy.json 2> gpurun_out/bench_galaxy.err; echo "rc=$?"; cut -c1-3000 gpurun_out/bench_galaxy.json; tail -3 gpurun_out/bench_galaxy.err
echo "== done (t=$SECONDS)"
