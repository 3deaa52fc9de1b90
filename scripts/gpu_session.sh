#!/bin/bash
# One GPU-box session: tests, ncu captures, benches.  Run as: gpurun --timeout 2400 -- 'bash scripts/gpu_session.sh'
mkdir -p gpurun_out
timeout 1400 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_gpu.log
grep -E "^E  " gpurun_out/pytest_gpu.log | grep -v "where\|array(" | head
ncu --set full --clock-control none --import-source on -k regex:direct_f32 -s 1 -c 1 -f -o gpurun_out/prof_direct_f32 python scripts/profile_kernels.py direct32 262144 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
ncu --set full --clock-control none --import-source on -k regex:direct_f64 -s 1 -c 1 -f -o gpurun_out/prof_direct_f64 python scripts/profile_kernels.py direct64 131072 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f -o gpurun_out/prof_walk_f32 python scripts/profile_kernels.py tree32 1048576 > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tree32.csv python scripts/profile_kernels.py tree32 4194304 > gpurun_out/ncu4.log 2>&1; tail -2 gpurun_out/ncu4.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu5.log 2>&1; tail -c 300 gpurun_out/ncu5.log
timeout 300 python bench.py --workload tree > gpurun_out/bench_tree.json 2> gpurun_out/bench_tree.err; cat gpurun_out/bench_tree.json; tail -3 gpurun_out/bench_tree.err
