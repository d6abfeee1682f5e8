#!/bin/bash
# GPU box: whole GPU suite, bench, ncu of the E-step kernel of the GEM loop (profile_driver = the bench model)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E   " gpurun_out/pytest_gpu.log | cut -c1-220 | tail -15
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python -c "
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e ms/step %.3f launches %d ll %.6f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['loglik_last']))
print('phases', d['phases_ms']); print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch']); print('clocks', d['clocks'])
"
CMD="python scripts/profile_driver.py --iters 3"
$CMD > gpurun_out/driver_plain.log 2>&1 || { echo plain failed; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:estep_stereo_kernel -s 2 -c 1 -f -o gpurun_out/r02_prof_estep_stereo_inst $CMD > gpurun_out/ncu_inst.log 2>&1; echo "ncu $?"
