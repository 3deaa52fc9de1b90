#!/bin/bash
# 1/2/4/8-GPU scaling of both workloads (run under gpurun --gpus 8)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for wl in direct tree; do
for n in 8 4 2 1; do
  if [ $n = 1 ]; then
    python bench.py --steps 5 --warmup 3 --workload $wl --no-cpu-baseline 2>gpurun_out/scale_${wl}_g$n.err | tail -1 > gpurun_out/scale_${wl}_g$n.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 5 --warmup 3 --workload $wl 2>gpurun_out/scale_${wl}_g$n.err | tail -1 > gpurun_out/scale_${wl}_g$n.json
  fi
  python -c "
import json
try:
    d=json.load(open('gpurun_out/scale_${wl}_g$n.json')); print('$wl', d['n_gpus'], '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], 'kernel %.3f ms'%d['roofline']['kernel_ms'], d.get('clocks'))
except Exception as e:
    print('$wl $n FAILED', e)
"
  tail -2 gpurun_out/scale_${wl}_g$n.err | cut -c1-300
done
done
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
