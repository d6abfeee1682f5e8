#!/bin/bash
# GPU box: multi-channel tests, then timing of the general-I E-step at two occupancies
mkdir -p gpurun_out
python -m pytest tests/test_multichannel_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout=900 > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E   " gpurun_out/pytest_multi.log | cut -c1-220 | tail -25
for lib in "" pyfasst_b200/libpyfasst_b200_minb3.so; do
  [ -n "$lib" ] && export PYFASST_B200_LIB=$lib
  python scripts/time_estep.py --I 4 --conv --rank 4 2>&1 | tail -1
  python scripts/time_estep.py --I 4 --rank 2 --dtype float64 2>&1 | tail -1
done
unset PYFASST_B200_LIB
python scripts/time_estep.py --I 3 --rank 2 --J 3 2>&1 | tail -1
PYFASST_FORCE_MULTI=1 python scripts/time_estep.py --I 2 --rank 2 2>&1 | tail -1
python scripts/time_estep.py 2>&1 | tail -1
