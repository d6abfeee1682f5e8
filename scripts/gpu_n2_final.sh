#!/bin/bash
# 2 GPUs, the driver's own launch line: the bench workload with the frames sharded (weak scaling)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_n2.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02b_bench_n2.json').read().strip().splitlines()[-1])
    print('n2: value %.4e e2e %.4e step %.3f ll %.9f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['loglik_last']), {k: round(v,3) for k,v in d['phases_ms'].items()})
    print('   ', {k: round(v,4) for k,v in d['e2e']['stages'].items()})
except Exception as e:
    print('n2 failed', e); print(open('gpurun_out/r02b_n2.err').read()[-1500:])
PY
