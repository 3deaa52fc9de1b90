#!/bin/bash
# round 2, session C: full GPU suite on one B200 after the phase / capacity rewrite of the tree, smoke, default bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/c_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/c_pytest.log
grep -E "^E  |^FAILED" gpurun_out/c_pytest.log | grep -v "where\|array(" | head -20
timeout 600 python bench.py > gpurun_out/c_bench_default.json 2> gpurun_out/c_bench_default.err; echo bench rc=$?; tail -3 gpurun_out/c_bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/c_launches_tree32.csv python scripts/profile_kernels.py tree32 4194304 > gpurun_out/c_ncu.log 2>&1; tail -1 gpurun_out/c_ncu.log
