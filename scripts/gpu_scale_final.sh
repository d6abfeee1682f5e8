#!/bin/bash
# weak-scaling points of the final state: bench.py --gpus N (frames sharded), N = all GPUs of the box
mkdir -p gpurun_out
NG=${NG:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/r01_bench_n${NG}_final.json 2> gpurun_out/n${NG}_final.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r01_bench_n${NG}_final.json').read().strip().splitlines()[-1])
    print('N=$NG: value %.4e e2e %.4e step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['phases_ms'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/n${NG}_final.err').read()[-1500:])
PY
