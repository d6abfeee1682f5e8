#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q -x --timeout=120 -k "tw or engine or gem" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_twq.json 2> gpurun_out/bench_twq.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_twq.json').read().strip().splitlines()[-1])
print('step %.3f ms  spectral %.3f  powers %.3f  estep %.3f  ll %.9f' % (d['ms_per_step'], d['phases_ms']['spectral'], d['phases_ms']['powers'], d['phases_ms']['estep'], d['loglik_last']))
PY
CMD="python scripts/profile_driver.py --iters 2"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:tw_contract_fused -s 1 -c 2 --csv $CMD 2>/dev/null | tail -4
