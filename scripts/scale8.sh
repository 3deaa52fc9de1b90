#!/bin/bash
mkdir -p gpurun_out
for wl in tree; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --steps 5 --warmup 3 --workload $wl 2>gpurun_out/scale_${wl}_g8.err | tail -1 > gpurun_out/scale_${wl}_g8.json
  python -c "
import json
d=json.load(open('gpurun_out/scale_${wl}_g8.json')); print('$wl', d['n_gpus'], '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], 'kernel %.3f ms'%d['roofline']['kernel_ms'], 'e2e %.4g'%d['e2e']['value'])
"
done
