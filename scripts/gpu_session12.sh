#!/bin/bash
mkdir -p gpurun_out
timeout 40 python bench.py --workload tree > gpurun_out/bench_tree.json 2> gpurun_out/bench_tree.err; echo "tree rc=$?"
timeout 60 python bench.py > gpurun_out/bench_direct.json 2> gpurun_out/bench_direct.err; echo "direct rc=$?"
python - <<'PY'
import json
for w in ("direct", "tree"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % w))
        print(w, "ms/step %.3f value %.4g e2e ms %.2f e2e value %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_call"], d["e2e"]["value"]))
    except Exception as e:
        print(w, "no result", e)
PY
