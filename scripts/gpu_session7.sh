#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:walk_group -s 1 -c 1 -f -o gpurun_out/prof_walk_group python scripts/profile_kernels.py tree32 4194304 > gpurun_out/ncu8.log 2>&1; tail -2 gpurun_out/ncu8.log
