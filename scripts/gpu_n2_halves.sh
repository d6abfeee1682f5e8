for h in 1 0; do
  PYFASST_TW_HALVES=$h timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --shard time > gpurun_out/n2_time_h$h.json 2> gpurun_out/n2.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/n2_time_h$h.json').read().strip().splitlines()[-1])
    print('halves=$h: value %.4e e2e %.4e step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v, 3) for k, v in d['phases_ms'].items()})
    print('   ', {k: round(v,4) for k,v in d['e2e']['stages'].items()})
except Exception as e:
    print('failed', e); print(open('gpurun_out/n2.err').read()[-800:])
PY
done
