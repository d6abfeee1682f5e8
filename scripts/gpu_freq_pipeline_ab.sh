#!/bin/bash
# 2 GPUs: frequency partition with / without the per-component pipeline (4-channel convolutive model)
mkdir -p gpurun_out
for p in 1 0; do
  PYFASST_FREQ_PIPELINE=$p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --model conv --channels 4 --rank 4 --duration-s 450 --shard freq > gpurun_out/freq_pipe_$p.json 2> gpurun_out/freq_pipe_$p.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/freq_pipe_$p.json').read().strip().splitlines()[-1])
    print('pipeline=$p: step %.3f ms ll %.9f' % (d['ms_per_step'], d['loglik_last']), d['phases_ms'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/freq_pipe_$p.err').read()[-1500:])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/trace_iteration.py --shard freq --channels 4 --model conv --rank 4 --duration-s 450 --out gpurun_out/trace_n2_4ch_freq_pipeline.txt 2>&1 | tail -1
