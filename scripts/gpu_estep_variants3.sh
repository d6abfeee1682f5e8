#!/bin/bash
# GPU box: what bounds the E-step?  float32 algebra (measurement only), interleaved splits
# (variant 35) and the number of passes per CTA.
mkdir -p gpurun_out
run() {  # label
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_x.json').read().strip().splitlines()[-1])
    print('$1: estep %.4f ms frac %.3f step %.3f ll %.9f' % (d['phases_ms']['estep'], d['roofline']['frac'], d['ms_per_step'], d['loglik_last']))
except Exception as e:
    print('$1 failed', e); print(open('gpurun_out/bench_x.err').read()[-800:])
PY
}
PYFASST_ESTEP_VARIANT=35 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout=300 2>&1 | tail -2
PYFASST_ESTEP_VARIANT=35 PYFASST_ESTEP_PASSES=4 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout=300 -k estep 2>&1 | tail -2
PYFASST_ESTEP_FLOAT_ALGEBRA=1 run "variant 3, float32 algebra (inaccurate; measurement only)"
for p in 32 8 4 2; do
  PYFASST_ESTEP_VARIANT=35 PYFASST_ESTEP_PASSES=$p run "variant 35 (interleaved), $p passes per CTA"
done
for p in 8 4; do
  PYFASST_ESTEP_VARIANT=3 PYFASST_ESTEP_PASSES=$p run "variant 3 (contiguous), $p passes per CTA"
done
