#!/bin/bash
# GPU box: interleaved tile / split mapping of spec_power_tc_kernel and fb_contract_tc_kernel
# (PYFASST_SPT_INTERLEAVE / PYFASST_FBT_INTERLEAVE, default on): parity, then phase times.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_nmf_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout=300 2>&1 | tail -2
run() {
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_x.json').read().strip().splitlines()[-1])
    print('$1: step %.3f ms  ll %.9f' % (d['ms_per_step'], d['loglik_last']), {k: round(v, 4) for k, v in d['phases_ms'].items()})
except Exception as e:
    print('$1 failed', e); print(open('gpurun_out/bench_x.err').read()[-800:])
PY
}
run "spec_power interleaved, fb_contract interleaved (default)"
PYFASST_SPT_INTERLEAVE=0 run "spec_power contiguous,  fb_contract interleaved"
PYFASST_FBT_INTERLEAVE=0 run "spec_power interleaved, fb_contract contiguous"
