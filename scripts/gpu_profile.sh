#!/bin/bash
# GPU box: full GPU test-suite (no -x), then bench + ncu launch list of the same command.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -30
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value %.3e e2e %.3e ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
print('phases', d['phases_ms']); print('roofline', d['roofline']); print('clocks', d['clocks']); print('cpu', d['cpu_baseline'])
"
PYFASST_ESTEP_FLOAT_ALGEBRA=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_floatalg.log 2>&1
python -c "
import json
d=json.loads(open('gpurun_out/bench_floatalg.log').read().strip().splitlines()[-1])
print('FLOAT-ALGEBRA estep: phases', d['phases_ms'])
"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_f64.log 2>&1
python -c "
import json
d=json.loads(open('gpurun_out/bench_f64.log').read().strip().splitlines()[-1])
print('F64: value %.3e phases'%d['value'], d['phases_ms'])
"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/bench_short.log 2> gpurun_out/bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu exit $?"
