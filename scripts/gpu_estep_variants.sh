#!/bin/bash
# GPU box: E-step tuning variants (PYFASST_ESTEP_VARIANT = OPT + 4 * (MINB - 2)): parity tests with
# the candidate, then the bench's E-step phase time for each variant.
mkdir -p gpurun_out
for v in ${TEST_VARIANTS:-3}; do
  PYFASST_ESTEP_VARIANT=$v python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_api_gpu.py -m gpu -q -x --timeout=600 2>&1 | tail -3
done
for v in ${VARIANTS:-0 1 2 3 4 5 6 7}; do
  PYFASST_ESTEP_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_v$v.json').read().strip().splitlines()[-1])
print('variant $v: estep %.4f ms  frac %.3f  step %.3f ms  ll %.9f' % (d['phases_ms']['estep'], d['roofline']['frac'], d['ms_per_step'], d['loglik_last']))
PY
done
