#!/bin/bash
# GPU box: the TW contraction with P' formed in the kernel (PYFASST_TW_FUSED=1): parity, then bench.
mkdir -p gpurun_out
PYFASST_TW_FUSED=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=120 -k "tw or spectral or nmf" 2>&1 | tail -6
PYFASST_TW_FUSED=1 timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_api_gpu.py -m gpu -q -x --timeout=300 2>&1 | tail -4
for fused in 0 1; do
  PYFASST_TW_FUSED=$fused timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_twf$fused.json 2> gpurun_out/bench_twf$fused.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_twf$fused.json').read().strip().splitlines()[-1])
    print('fused=$fused: step %.3f ms  spectral %.3f  powers %.3f  estep %.3f  ll %.9f  launches %d' % (d['ms_per_step'], d['phases_ms']['spectral'], d['phases_ms']['powers'], d['phases_ms']['estep'], d['loglik_last'], d['gpu_launches']))
except Exception as e:
    print('fused=$fused failed', e); print(open('gpurun_out/bench_twf$fused.err').read()[-1500:])
PY
done
