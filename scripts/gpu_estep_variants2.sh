#!/bin/bash
# GPU box: E-step variants by number (PYFASST_ESTEP_VARIANT): parity (kernel, engine, API, full-size
# tests), then the bench's E-step phase time.
mkdir -p gpurun_out
for v in ${VARIANTS:-3 19}; do
  export PYFASST_ESTEP_VARIANT=$v
  timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_api_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout=300 2>&1 | tail -4
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_v$v.json').read().strip().splitlines()[-1])
    print('variant $v: estep %.4f ms frac %.3f step %.3f ll %.9f' % (d['phases_ms']['estep'], d['roofline']['frac'], d['ms_per_step'], d['loglik_last']))
except Exception as e:
    print('variant $v failed', e); print(open('gpurun_out/bench_v$v.err').read()[-800:])
PY
done
