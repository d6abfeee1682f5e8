#!/bin/bash
# GPU box: ncu --set full captures of the three big kernels (one launch each).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -10
CMD="python scripts/profile_driver.py --iters 2"
$CMD > gpurun_out/driver_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/driver_plain.log; exit 1; }
for kern in estep_stereo_kernel fb_contract_same_kernel tw_contract_kernel spec_power_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$kern -s 4 -c 1 \
      -f -o gpurun_out/prof_$kern $CMD > gpurun_out/ncu_$kern.log 2>&1
  echo "ncu $kern exit $?"
done
ls -la gpurun_out/*.ncu-rep
