#!/bin/bash
mkdir -p gpurun_out
timeout 1400 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu.log
grep -E "^E  " gpurun_out/pytest_gpu.log | grep -v "where\|array(" | head
timeout 300 python bench.py --workload tree --no-cpu-baseline > gpurun_out/bench_tree.json 2> gpurun_out/bench_tree.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tree.json')); print('tree', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_call'])"; tail -3 gpurun_out/bench_tree.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tree32.csv python scripts/profile_kernels.py tree32 4194304 > gpurun_out/ncu4.log 2>&1; tail -1 gpurun_out/ncu4.log
ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f -o gpurun_out/prof_walk_f32 python scripts/profile_kernels.py tree32 4194304 > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log
python scripts/gpu_probe.py 2>&1 | grep -E "^tree|stats" | tail -12
