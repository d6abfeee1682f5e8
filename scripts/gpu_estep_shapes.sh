#!/bin/bash
# GPU box: E-step CTA shapes (threads x __maxnreg__) built as variant libraries.
mkdir -p gpurun_out
for lib in "" pyfasst_b200/variants/lib_estep_96_224.so pyfasst_b200/variants/lib_estep_160_200.so pyfasst_b200/variants/lib_estep_64_200.so; do
  export PYFASST_B200_LIB=$lib
  [ -n "$lib" ] && export PYFASST_B200_LIB=$PWD/$lib
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=120 -k "test_estep_stereo and not warp" 2>&1 | tail -1
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_shape.json 2> gpurun_out/bench_shape.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_shape.json').read().strip().splitlines()[-1])
    print('${lib:-default}: estep %.4f ms frac %.3f step %.3f ll %.9f' % (d['phases_ms']['estep'], d['roofline']['frac'], d['ms_per_step'], d['loglik_last']))
except Exception as e:
    print('${lib:-default} failed', e); print(open('gpurun_out/bench_shape.err').read()[-800:])
PY
done
