#!/bin/bash
# A/B of group-walk register caps (scripts/build_variants.py) on the tree bench, N = 4M.
mkdir -p gpurun_out
V=gravhopper_b200/variants
one() {  # name lib block
  local name=$1 lib=$2 blk=$3
  GH_B200_LIB=$lib GH_WALK_BLOCK=$blk timeout 120 python bench.py --workload tree --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err || { echo "$name FAILED"; tail -3 gpurun_out/var_$name.err; return; }
  python - "$name" <<'PY'
import json, sys
d = json.load(open("gpurun_out/var_%s.json" % sys.argv[1]))
r = d["roofline"]
print("%-10s ms/step %.3f walk %.3f build %.3f list %.1f err mean %.6e max %.6e" % (sys.argv[1], d["ms_per_step"], r["kernel_ms"], r["build_ms"], r["accepted_per_target"], r["accuracy"]["timed_fp32_walk"]["mean"], r["accuracy"]["timed_fp32_walk"]["max"]))
PY
}
one base64 "" 64
one w24_64 $PWD/$V/lib_w24.so 64
one w20_64 $PWD/$V/lib_w20.so 64
one w16_64 $PWD/$V/lib_w16.so 64
one w20_128 $PWD/$V/lib_w20.so 128
one w16_128 $PWD/$V/lib_w16.so 128
one w16_32 $PWD/$V/lib_w16.so 32
one base64b "" 64
echo "done t=$SECONDS"
