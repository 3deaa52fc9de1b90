#!/bin/bash
mkdir -p gpurun_out
timeout 1400 python -m pytest tests -q -m gpu -x -k "tree or sharded or first_ten or kats or ragged" > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu.log
grep -E "^E  " gpurun_out/pytest_gpu.log | grep -v "where\|array(" | head
timeout 300 python bench.py --workload tree --no-cpu-baseline > gpurun_out/bench_tree.json 2> gpurun_out/bench_tree.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tree.json')); print('tree', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_call'])"; tail -3 gpurun_out/bench_tree.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tree32.csv python scripts/profile_kernels.py tree32 4194304 > gpurun_out/ncu4.log 2>&1; tail -1 gpurun_out/ncu4.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tree64.csv python scripts/profile_kernels.py tree64 4194304 > gpurun_out/ncu6.log 2>&1; tail -1 gpurun_out/ncu6.log
