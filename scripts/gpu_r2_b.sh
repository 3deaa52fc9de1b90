#!/bin/bash
# round 2, session B: hybrid kappa sweep over all 4M particles; default bench line with the tree block
mkdir -p gpurun_out
timeout 500 python scripts/gpu_hybrid_sweep.py 4194304 0 0.1 0.15 0.2 0.3 > gpurun_out/b_sweep.log 2>&1; tail -8 gpurun_out/b_sweep.log
timeout 600 python bench.py > gpurun_out/b_bench_default.json 2> gpurun_out/b_bench_default.err; echo bench rc=$?; cut -c1-600 gpurun_out/b_bench_default.json; tail -5 gpurun_out/b_bench_default.err
