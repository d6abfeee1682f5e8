#!/bin/bash
# GPU box: the instantaneous-mixing E-step: kernel + engine + real-size parity tests, timing, bench
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_tamy_gpu.py tests/test_api_gpu.py tests/test_fullsize_gpu.py tests/test_boundary_gpu.py tests/test_batch_gpu.py -m gpu -q --timeout=900 > gpurun_out/pytest_inst.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E   " gpurun_out/pytest_inst.log | cut -c1-220 | tail -15
python scripts/time_estep.py --for-update 2>&1 | tail -1
python scripts/time_estep.py 2>&1 | tail -1
python scripts/time_estep.py --for-update --dtype float64 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e ms/step %.3f launches %d ll %.6f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['loglik_last']))
print('phases', d['phases_ms']); print('roofline frac %.3f' % d['roofline']['frac']); print('clocks', d['clocks'])
"
