#!/bin/bash
# Second validation pass: full GPU suite with the model-parity tests, tree bench with and without
# page-locked result arrays (e2e A/B).
mkdir -p gpurun_out
BUDGET=${GH_SESSION_BUDGET:-400}
left() { echo $(( BUDGET - SECONDS )); }
run() { local need=$1 to=$2; shift 2; if [ $(left) -lt $need ]; then echo "SKIP: $*"; return 99; fi; [ $to -gt $(left) ] && to=$(left); timeout $to "$@"; }
echo "== pytest (t=$SECONDS)"
run 120 300 python -m pytest tests -q -m gpu -n 3 --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -22 gpurun_out/pytest_gpu.log | cut -c1-200
echo "== bench tree (t=$SECONDS)"
run 60 200 python bench.py --workload tree > gpurun_out/bench_tree.json 2> gpurun_out/bench_tree.err; echo "rc=$?"; tail -3 gpurun_out/bench_tree.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_tree.json",):
    d = json.load(open(f)); print(f, "ms/step", d["ms_per_step"], "value %.4g" % d["value"], "e2e", d["e2e"])
PY
echo "== bench tree, pageable results (t=$SECONDS)"
GH_PINNED_OUTPUT=0 run 60 200 python bench.py --workload tree --no-cpu-baseline > gpurun_out/bench_tree_pageable.json 2> gpurun_out/bench_tree_pageable.err; echo "rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_tree_pageable.json")); print("pageable e2e", d["e2e"])
PY
echo "== bench direct (t=$SECONDS)"
run 60 200 python bench.py --no-cpu-baseline > gpurun_out/bench_direct_nocpu.json 2> gpurun_out/bench_direct.err; echo "rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_direct_nocpu.json")); print("direct value %.4g" % d["value"], "e2e", d["e2e"], "frac", d["roofline"]["frac"])
PY
echo "== done (t=$SECONDS)"
