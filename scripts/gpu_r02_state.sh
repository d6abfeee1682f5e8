#!/bin/bash
# GPU box: state check of the round: GPU tests, bench, launch list, E-step timings (stereo + 4 channels)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|^E   " gpurun_out/pytest_gpu.log | cut -c1-200 | tail -30
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -5 gpurun_out/bench.err
python -c "
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e ms/step %.3f launches %d ll %.6f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['loglik_last']))
print('phases', d['phases_ms']); print('roofline frac %.3f' % d['roofline']['frac']); print('clocks', d['clocks'])
"
python scripts/time_estep.py 2>&1 | tail -1
python scripts/time_estep.py --conv 2>&1 | tail -1
python scripts/time_estep.py --dtype float64 2>&1 | tail -1
python scripts/time_estep.py --I 4 --conv --rank 4 2>&1 | tail -1
python scripts/time_estep.py --I 4 --rank 2 2>&1 | tail -1
