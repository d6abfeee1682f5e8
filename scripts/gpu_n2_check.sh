#!/bin/bash
# 2 GPUs: both partitions of the bench workload (weak scaling), kernel timeline of the frequency partition
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "pack_chunks" 2>&1 | tail -1
for shard in time freq; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline --shard $shard > gpurun_out/r02_bench_n2_${shard}.json 2> gpurun_out/n2_$shard.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n2_${shard}.json').read().strip().splitlines()[-1])
    print('$shard: value %.4e e2e %.4e step %.3f ll %.9f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['loglik_last']), d['phases_ms'])
    print('   ', {k: round(v,4) for k,v in d['e2e']['stages'].items()})
except Exception as e:
    print('$shard failed', e); print(open('gpurun_out/n2_$shard.err').read()[-1500:])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/trace_iteration.py --shard freq --out gpurun_out/trace_n2_freq.txt 2>&1 | tail -1
