#!/bin/bash
# A/B of library variants (scripts/build_variants.py) on the quick tree bench: scripts/gpu_ab.sh name1 name2 ...
# ("base" = the in-tree library).  Extra environment for every run: GH_AB_ENV="K=V K=V"; GH_AB_LABEL
# is appended to the output names (several environments of one library).
mkdir -p gpurun_out
V=$PWD/gravhopper_b200/variants
L=${GH_AB_LABEL:+_$GH_AB_LABEL}
for name in "$@"; do
  lib=$V/lib_$name.so; [ "$name" = base ] && lib=$PWD/gravhopper_b200/libgravhopper_b200.so
  env GH_B200_LIB=$lib $GH_AB_ENV timeout 200 python bench.py --workload tree --steps 30 --warmup 5 --quick > gpurun_out/ab_$name$L.json 2> gpurun_out/ab_$name$L.err || { echo "$name FAILED"; tail -3 gpurun_out/ab_$name$L.err; continue; }
  python - "$name$L" <<'PY'
import json, sys
d = json.loads([l for l in open("gpurun_out/ab_%s.json" % sys.argv[1]) if l.startswith("{")][-1]); r = d["roofline"]
print("%-10s ms/step %.3f walk %.3f build %.3f list %.1f err mean %.6e" % (sys.argv[1], d["ms_per_step"], r["kernel_ms"], r["build_ms"], r["accepted_per_target"], r["accuracy"]["timed_fp32_walk"]["mean"]))
PY
done
