#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_gpu.log
grep -E "^E  |^FAILED" gpurun_out/pytest_gpu.log | grep -v "where\|array(" | head
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_driver.py 3000 > gpurun_out/memcheck.log 2>&1; echo memcheck rc=$?; tail -2 gpurun_out/memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_driver.py 2500 > gpurun_out/racecheck.log 2>&1; echo racecheck rc=$?; tail -2 gpurun_out/racecheck.log
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 3 python scripts/sanitize_driver.py 2500 > gpurun_out/synccheck.log 2>&1; echo synccheck rc=$?; tail -2 gpurun_out/synccheck.log
timeout 400 compute-sanitizer --tool initcheck --error-exitcode 3 python scripts/sanitize_driver.py 2500 > gpurun_out/initcheck.log 2>&1; echo initcheck rc=$?; tail -2 gpurun_out/initcheck.log
