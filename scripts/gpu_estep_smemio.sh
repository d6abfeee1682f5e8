#!/bin/bash
# GPU box: E-step with shared-memory-resident I/O (PYFASST_ESTEP_VARIANT=15) at several CTA shapes
# (variant libraries built with -DPF_ESTEP_THREADS / -DPF_ESTEP_MAXNREG): parity, then the bench's
# E-step phase time.
mkdir -p gpurun_out
run() {  # label, lib, variant
  export PYFASST_B200_LIB=$2
  [ -n "$2" ] && export PYFASST_B200_LIB=$PWD/$2
  export PYFASST_ESTEP_VARIANT=$3
  timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x --timeout=120 -k "estep and not warp" 2>&1 | tail -1
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_shape.json 2> gpurun_out/bench_shape.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_shape.json').read().strip().splitlines()[-1])
    print('$1: estep %.4f ms frac %.3f step %.3f ll %.9f' % (d['phases_ms']['estep'], d['roofline']['frac'], d['ms_per_step'], d['loglik_last']))
except Exception as e:
    print('$1 failed', e); print(open('gpurun_out/bench_shape.err').read()[-800:])
PY
}
run "default (128 thr, 255 regs, variant 3)" "" 3
run "smem I/O 128 thr x 168 regs (3 CTAs)" "" 15
for v in ${SHAPES:-160_200 64_200 96_224}; do
  run "smem I/O $v" pyfasst_b200/variants/lib_estep_$v.so 15
done
