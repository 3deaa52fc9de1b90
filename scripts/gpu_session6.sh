#!/bin/bash
# group-walk bring-up: accuracy/edge cases, then the tree bench in both walk modes
mkdir -p gpurun_out
timeout 400 python scripts/gpu_groupwalk.py > gpurun_out/groupwalk.log 2>&1; echo groupwalk rc=$?
tail -45 gpurun_out/groupwalk.log
for mode in target group; do
  GH_TREE_WALK=$mode timeout 300 python bench.py --workload tree --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tree_$mode.json 2> gpurun_out/bench_tree_$mode.err; echo bench $mode rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_tree_$mode.json"))
    print("$mode", "ms/step", d["ms_per_step"], "walk ms", d["roofline"]["kernel_ms"], "value", d["value"], "acc/target", d["roofline"]["accepted_per_target"])
except Exception as e:
    print("no bench json", e)
PY
done
