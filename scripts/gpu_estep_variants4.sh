#!/bin/bash
# GPU box: interleaved splits (variant 35) -- passes per CTA and the 256-thread CTA shape.
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
export PYFASST_ESTEP_VARIANT=35
for p in 64 16; do
  PYFASST_ESTEP_PASSES=$p run "variant 35, 128 threads, $p passes per CTA"
done
export PYFASST_B200_LIB=$PWD/pyfasst_b200/variants/lib_estep_t256.so
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout=300 -k estep 2>&1 | tail -2
for p in 32 16 8; do
  PYFASST_ESTEP_PASSES=$p run "variant 35, 256 threads x 1 CTA/SM, $p passes per CTA"
done
PYFASST_ESTEP_VARIANT=3 run "variant 3, 256 threads x 1 CTA/SM, 32 passes per CTA"
