#!/bin/bash
# plain-double moment scan (fp32 tree): tests, tree bench, launch list, ncu of the two largest build kernels
mkdir -p gpurun_out
echo "== pytest tree-related (t=$SECONDS)"
timeout 300 python -m pytest tests -q -m gpu -n 3 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-200
echo "== bench tree (t=$SECONDS)"
timeout 200 python bench.py --workload tree --steps 20 > gpurun_out/bench_tree.json 2> gpurun_out/bench_tree.err; echo "rc=$?"; tail -3 gpurun_out/bench_tree.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_tree.json")); r = d["roofline"]
print("ms/step", d["ms_per_step"], "walk", r["kernel_ms"], "build", r["build_ms"], "value %.4g" % d["value"], "e2e ms", d["e2e"]["ms_per_call"], r["accuracy"]["timed_fp32_walk"])
PY
echo "== launch list (t=$SECONDS)"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tree32.csv python scripts/profile_kernels.py tree32 4194304 > gpurun_out/ncu_tree.log 2>&1; tail -1 gpurun_out/ncu_tree.log
echo "== ncu full: scatter, emit (t=$SECONDS)"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:rs_scatter -s 12 -c 1 -f -o gpurun_out/prof_rs_scatter python scripts/profile_kernels.py tree32 4194304 > gpurun_out/ncu_sc.log 2>&1; tail -1 gpurun_out/ncu_sc.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:emit_kernel -s 1 -c 1 -f -o gpurun_out/prof_emit python scripts/profile_kernels.py tree32 4194304 > gpurun_out/ncu_em.log 2>&1; tail -1 gpurun_out/ncu_em.log
ls -la gpurun_out/*.ncu-rep
echo "== done (t=$SECONDS)"
