#!/bin/bash
# 2 GPUs, 4-channel convolutive model: hardware queue count (CUDA_DEVICE_MAX_CONNECTIONS) x partition
mkdir -p gpurun_out
for conn in 8 32; do for shard in time freq; do
  CUDA_DEVICE_MAX_CONNECTIONS=$conn timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --model conv --channels 4 --rank 4 --duration-s 112.5 --blocks 1 --shard $shard > gpurun_out/conn_${conn}_$shard.json 2> gpurun_out/conn.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/conn_${conn}_$shard.json').read().strip().splitlines()[-1])
    print('conn=$conn $shard: step %.3f ms' % d['ms_per_step'], {k: round(v, 3) for k, v in d['phases_ms'].items()})
except Exception as e:
    print('failed', e); print(open('gpurun_out/conn.err').read()[-800:])
PY
done; done
