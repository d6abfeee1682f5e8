#!/bin/bash
# GPU box: kernel + engine tests, then a short bench with phase times.
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
print('phases', d['phases_ms']); print('roofline frac %.3f' % d['roofline']['frac'])
"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
python scripts/launch_summary.py gpurun_out/launches.csv 2>/dev/null | head -12
