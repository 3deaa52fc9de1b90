#!/bin/bash
# round 2, session A: the two gated GPU tests, hybrid cost at N = 4M
mkdir -p gpurun_out
GH_TEST_HYBRID=1 GH_TEST_EXPDISK=1 timeout 600 python -m pytest tests/test_gpu_groupwalk.py tests/test_gpu_ic.py -q -m gpu -x > gpurun_out/a_pytest.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/a_pytest.log
timeout 300 python bench.py --workload tree --no-cpu-baseline > gpurun_out/a_tree_plain.json 2> gpurun_out/a_tree_plain.err; cut -c1-300 gpurun_out/a_tree_plain.json
GH_WALK_HYBRID=0.1 timeout 300 python bench.py --workload tree --no-cpu-baseline > gpurun_out/a_tree_hybrid.json 2> gpurun_out/a_tree_hybrid.err; cut -c1-300 gpurun_out/a_tree_hybrid.json
GH_WALK_HYBRID=0.2 timeout 300 python bench.py --workload tree --no-cpu-baseline > gpurun_out/a_tree_hybrid02.json 2> gpurun_out/a_tree_hybrid02.err; cut -c1-300 gpurun_out/a_tree_hybrid02.json
