#!/bin/bash
# GPU box: the warp-specialised E-step: parity, then the bench's E-step phase for each kernel / team layout.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=300 -k "estep" 2>&1 | tail -8
for cfg in ${CFGS:-"fused 0" "ws 0" "ws 1" "ws 2" "ws 3"}; do
  set -- $cfg
  PYFASST_ESTEP_KERNEL=$1 PYFASST_ESTEP_WSCFG=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1$2.json 2> gpurun_out/bench_$1$2.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$1$2.json').read().strip().splitlines()[-1])
    print('$1 cfg $2: estep %.4f ms  frac %.3f  step %.3f ms  ll %.9f' % (d['phases_ms']['estep'], d['roofline']['frac'], d['ms_per_step'], d['loglik_last']))
except Exception as e:
    print('$1 $2 failed', e); print(open('gpurun_out/bench_$1$2.err').read()[-1500:])
PY
done
