#!/bin/bash
# One parametrised GPU session script (run under gpurun; replaces the per-session one-offs).
#   scripts/gpu.sh <task> [<task> ...]       tasks run in order, each under its own timeout; outputs in gpurun_out/
# tasks:
#   smoke                      __graft_entry__.smoke()
#   tests[:<pytest args>]      pytest -m gpu -x -q [args]                  -> <tag>_pytest.log
#   bench[:<bench args>]       python bench.py [args]                      -> <tag>_bench_<n>.json
#   mbench:<N>[:<bench args>]  torchrun --nproc-per-node N bench.py --gpus N [args]
#   launches:<which>:<n>[:<steps>]  ncu launch list of scripts/profile_kernels.py <which> <n> [steps]
#   ab:<name,name,..>[:K=V,K=V[:label]]  scripts/gpu_ab.sh: quick tree bench with library variants (base = in-tree)
#   ncu:<kernel-regex>:<which>:<n>[:<skip>]  ncu --set full of one launch  -> <tag>_<regex>.ncu-rep
#   benchlaunches[:<bench args>]  ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (a number
#                              printed under ncu is never a bench value)
#   sweep[:<N>[:kappas...]]    scripts/gpu_hybrid_sweep.py
#   py:<script and args>       python <script> ...
# GH_TAG names the outputs (default r02).  Multi-GPU benches run under a short timeout of their own
# (GH_MBENCH_TIMEOUT, default 300 s): a hung collective must not hold N GPUs until gpurun's limit.
mkdir -p gpurun_out
TAG=${GH_TAG:-r02}
k=0
for task in "$@"; do
  k=$((k+1))
  IFS=':' read -r what a1 a2 a3 a4 <<< "$task"
  echo "== [$k] $task (t=${SECONDS}s)"
  case "$what" in
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
    tests) timeout 2400 python -m pytest tests -q -m gpu -x $a1 > gpurun_out/${TAG}_pytest_$k.log 2>&1; echo "pytest rc=$?"
           tail -4 gpurun_out/${TAG}_pytest_$k.log | cut -c1-240
           grep -E "^E  |^FAILED" gpurun_out/${TAG}_pytest_$k.log | grep -v "where\|array(" | head -20 ;;
    bench) timeout 900 python bench.py $a1 > gpurun_out/${TAG}_bench_$k.json 2> gpurun_out/${TAG}_bench_$k.err; echo "bench rc=$?"
           tail -3 gpurun_out/${TAG}_bench_$k.err; cut -c1-300 gpurun_out/${TAG}_bench_$k.json ;;
    mbench) timeout ${GH_MBENCH_TIMEOUT:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $a1 --master-addr 127.0.0.1 --master-port $((29600+k)) \
              bench.py --gpus $a1 $a2 > gpurun_out/${TAG}_mbench${a1}_$k.json 2> gpurun_out/${TAG}_mbench${a1}_$k.err; echo "bench rc=$?"
           tail -3 gpurun_out/${TAG}_mbench${a1}_$k.err | cut -c1-300; tail -1 gpurun_out/${TAG}_mbench${a1}_$k.json | cut -c1-300 ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_${a1}_${a2}.csv \
              python scripts/profile_kernels.py $a1 $a2 $a3 > gpurun_out/${TAG}_ncu_$k.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_$k.log ;;
    ab) GH_AB_ENV="${a2//,/ }" GH_AB_LABEL="$a3" bash scripts/gpu_ab.sh ${a1//,/ } ;;
    ncu) timeout 900 ncu --set full --clock-control none --import-source on -k regex:$a1 -s ${a4:-1} -c 1 -f -o gpurun_out/${TAG}_${a1}_${a3} \
              python scripts/profile_kernels.py $a2 $a3 4 > gpurun_out/${TAG}_ncu_$k.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_$k.log ;;
    benchlaunches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
              python bench.py --steps 2 --warmup 3 --no-cpu-baseline $a1 > gpurun_out/${TAG}_ncu_$k.log 2>&1; tail -c 300 gpurun_out/${TAG}_ncu_$k.log ;;
    sweep) timeout 600 python scripts/gpu_hybrid_sweep.py ${a1:-4194304} ${a2//,/ } > gpurun_out/${TAG}_sweep_$k.log 2>&1; tail -8 gpurun_out/${TAG}_sweep_$k.log | cut -c1-300 ;;
    py) timeout 900 python $a1 > gpurun_out/${TAG}_py_$k.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/${TAG}_py_$k.log | cut -c1-300 ;;
    *) echo "unknown task $what" ;;
  esac
done
echo "== done (t=${SECONDS}s)"
