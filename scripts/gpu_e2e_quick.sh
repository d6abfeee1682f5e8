#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_api_gpu.py tests/test_engine_gpu.py tests/test_batch_gpu.py tests/test_sourcefilter_gpu.py tests/test_multichannel_gpu.py -m gpu -q -x --timeout=300 2>&1 | tail -2
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v*1e3,2) for k,v in d['e2e']['stages'].items()})
PY
done
