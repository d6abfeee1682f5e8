#!/bin/bash
# GPU box: E-step kernel tests, then A/B timing of the lane-pair kernel against the one-thread-per-bin kernel
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_tamy_gpu.py tests/test_engine_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x --timeout=900 -k "${KEXPR:-estep or tamy or engine or fullsize}" > gpurun_out/pytest_estep.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E   " gpurun_out/pytest_estep.log | cut -c1-200 | tail -15
for v in 1 0; do
  echo "PYFASST_ESTEP_PAIR=$v"
  PYFASST_ESTEP_PAIR=$v python scripts/time_estep.py 2>&1 | tail -1
  PYFASST_ESTEP_PAIR=$v python scripts/time_estep.py --dtype float64 2>&1 | tail -1
done
for p in 16 32 128; do echo "passes $p"; PYFASST_ESTEP_PASSES=$p python scripts/time_estep.py 2>&1 | tail -1; done
