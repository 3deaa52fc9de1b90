#!/bin/bash
# group-walk bring-up: accuracy/edge cases, then the tree bench in both walk modes / CTA sizes
mkdir -p gpurun_out
timeout 400 python scripts/gpu_groupwalk.py > gpurun_out/groupwalk.log 2>&1; echo groupwalk rc=$?
grep -E "group|determ" gpurun_out/groupwalk.log | grep -v "^n=" | tail -30
grep -E "^n=" gpurun_out/groupwalk.log | grep group
for wb in 64 128; do
  GH_WALK_BLOCK=$wb GH_TREE_WALK=group timeout 300 python bench.py --workload tree --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tree_group$wb.json 2> gpurun_out/bench_tree_group$wb.err; echo bench $wb rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_tree_group$wb.json"))
    print("block $wb", "ms/step", d["ms_per_step"], "walk ms", d["roofline"]["kernel_ms"], "value", d["value"], "acc/target", d["roofline"]["accepted_per_target"])
except Exception as e:
    print("no bench json", e)
PY
done
GH_TREE_WALK=target timeout 300 python bench.py --workload tree --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tree_target.json 2> gpurun_out/bench_tree_target.err; echo bench target rc=$?
python -c "
import json
d = json.load(open('gpurun_out/bench_tree_target.json'))
print('target', 'ms/step', d['ms_per_step'], 'walk ms', d['roofline']['kernel_ms'], 'value', d['value'], 'acc/target', d['roofline']['accepted_per_target'])
"
