#!/bin/bash
# Final validation of the round: serial GPU suite exactly as the driver runs it, smoke, default /
# tree / galaxy bench lines for profiles/, then (time permitting) the emit-kernel occupancy A/B.
mkdir -p gpurun_out
BUDGET=${GH_SESSION_BUDGET:-320}
left() { echo $(( BUDGET - SECONDS )); }
run() { local need=$1 to=$2; shift 2; if [ $(left) -lt $need ]; then echo "SKIP: $*"; return 99; fi; [ $to -gt $(left) ] && to=$(left); timeout $to "$@"; }
echo "== pytest serial -x (t=$SECONDS)"
run 100 200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-200
echo "== smoke (t=$SECONDS)"
run 30 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench direct (t=$SECONDS)"
run 50 120 python bench.py > gpurun_out/bench_direct.json 2> gpurun_out/bench_direct.err; echo "rc=$?"
echo "== bench tree (t=$SECONDS)"
run 30 100 python bench.py --workload tree > gpurun_out/bench_tree.json 2> gpurun_out/bench_tree.err; echo "rc=$?"
echo "== bench galaxy (t=$SECONDS)"
run 40 100 python bench.py --workload galaxy --no-cpu-baseline > gpurun_out/bench_galaxy.json 2> gpurun_out/bench_galaxy.err; echo "rc=$?"
python - <<'PY'
import json
for w in ("direct", "tree", "galaxy"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % w)); r = d["roofline"]
        print(w, "ms/step %.3f value %.4g kernel_ms %.3f frac %.3f e2e ms %.2f launches %d" % (d["ms_per_step"], d["value"], r["kernel_ms"], r["frac"], d["e2e"]["ms_per_call"], d["gpu_launches"]), r.get("accuracy", {}).get("timed_fp32_walk"))
    except Exception as e:
        print(w, "no result", e)
PY
V=gravhopper_b200/variants
one() {
  local name=$1 lib=$2
  GH_B200_LIB=$lib run 25 60 python -m pytest tests/test_gpu_parity.py -q -k "tree or ragged" > gpurun_out/var_$name.test 2>&1; local trc=$?
  GH_B200_LIB=$lib run 20 60 python bench.py --workload tree --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err || { echo "$name skipped/failed"; return; }
  python - "$name" "$trc" <<'PY'
import json, sys
d = json.load(open("gpurun_out/var_%s.json" % sys.argv[1])); r = d["roofline"]
print("%-8s tests rc=%s ms/step %.3f walk %.3f build %.3f err mean %.6e" % (sys.argv[1], sys.argv[2], d["ms_per_step"], r["kernel_ms"], r["build_ms"], r["accuracy"]["timed_fp32_walk"]["mean"]))
PY
}
for v in emit12 emit10 emit16 emit8; do one $v $PWD/$V/lib_$v.so; done
echo "== done (t=$SECONDS)"
