#!/bin/bash
# A/B of radix-sort scatter variants (scripts/build_variants.py) on the tree bench, N = 4M; every
# variant first has to pass the tree parity tests (the fp64 tree is bit-sensitive to the sort).
mkdir -p gpurun_out
V=gravhopper_b200/variants
one() {
  local name=$1 lib=$2
  GH_B200_LIB=$lib timeout 120 python -m pytest tests/test_gpu_parity.py -q -k "tree or ragged" > gpurun_out/var_$name.test 2>&1; local trc=$?
  GH_B200_LIB=$lib timeout 120 python bench.py --workload tree --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err || { echo "$name bench FAILED"; tail -3 gpurun_out/var_$name.err; return; }
  python - "$name" "$trc" <<'PY'
import json, sys
d = json.load(open("gpurun_out/var_%s.json" % sys.argv[1]))
r = d["roofline"]
print("%-8s tests rc=%s ms/step %.3f walk %.3f build %.3f err mean %.6e" % (sys.argv[1], sys.argv[2], d["ms_per_step"], r["kernel_ms"], r["build_ms"], r["accuracy"]["timed_fp32_walk"]["mean"]))
PY
}
one base ""
for v in rs1 rs2 rs2b3 rs2b5 rs0b3 rs0b5; do one $v $PWD/$V/lib_$v.so; done
one base2 ""
echo "done t=$SECONDS"
