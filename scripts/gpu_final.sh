#!/bin/bash
# Final validation on one B200: smoke, full GPU tests, default bench (+ reference arm), refreshed ncu evidence.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_gpu.log
grep -E "^E  |^FAILED" gpurun_out/pytest_gpu.log | grep -v "where\|array(" | head
python bench.py --impl reference > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err; cut -c1-260 gpurun_out/bench_reference.json
python bench.py > gpurun_out/bench_direct.json 2>gpurun_out/bench_direct.err; cat gpurun_out/bench_direct.json; tail -2 gpurun_out/bench_direct.err
python bench.py --workload tree > gpurun_out/bench_tree.json 2>gpurun_out/bench_tree.err; cut -c1-400 gpurun_out/bench_tree.json; tail -2 gpurun_out/bench_tree.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu5.log 2>&1; tail -c 200 gpurun_out/ncu5.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tree32.csv python scripts/profile_kernels.py tree32 4194304 > gpurun_out/ncu4.log 2>&1; tail -1 gpurun_out/ncu4.log
ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f -o gpurun_out/prof_walk_f32 python scripts/profile_kernels.py tree32 4194304 > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log
ncu --set full --clock-control none --import-source on -k regex:direct_f64 -s 1 -c 1 -f -o gpurun_out/prof_direct_f64 python scripts/profile_kernels.py direct64 131072 > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
ncu --set full --clock-control none --import-source on -k regex:direct_f32 -s 2 -c 1 -f -o gpurun_out/prof_direct_f32_N1M python scripts/profile_kernels.py direct32 1048576 > gpurun_out/ncu7.log 2>&1; tail -1 gpurun_out/ncu7.log
