#!/bin/bash
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
grep -E "^E  " gpurun_out/pytest_gpu.log | grep -v "where\|array(" | head
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_driver.py 1200 > gpurun_out/memcheck.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_driver.py 700 > gpurun_out/racecheck.log 2>&1; echo racecheck rc=$?; tail -3 gpurun_out/racecheck.log
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 3 python scripts/sanitize_driver.py 700 > gpurun_out/synccheck.log 2>&1; echo synccheck rc=$?; tail -2 gpurun_out/synccheck.log
python scripts/gpu_probe.py 2>&1 | grep -E "^direct fp64 (65536|262144)" 
ncu --set full --clock-control none --import-source on -k regex:direct_f32 -s 2 -c 1 -f -o gpurun_out/prof_direct_f32_N1M python scripts/profile_kernels.py direct32 1048576 > gpurun_out/ncu7.log 2>&1; tail -1 gpurun_out/ncu7.log
