#!/bin/bash
# GPU box: compute-sanitizer memcheck / racecheck over the small-fixture kernel tests.
# NOTE (round 2): this GPU pool refuses compute-sanitizer ("closed on this pool and stays closed: runs
# under it have left GPUs needing a reset") -- profiles/r02/sanitizer_closed_on_this_pool.log is the
# output of this very command.  On a pool that allows it, the logs land in gpurun_out/.
TAG=${TAG:-r02}
mkdir -p gpurun_out
SAN="compute-sanitizer --error-exitcode 86 --launch-timeout 0"
timeout 1500 $SAN --tool memcheck python -m pytest tests/test_kernels_gpu.py tests/test_multichannel_gpu.py -m gpu -q -x --timeout=1400 \
    -k "estep or wiener or contract or spec_power or mix or stft" > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; tail -4 gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 1500 $SAN --tool racecheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=1400 \
    -k "estep or contract or spec_power" > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?"; tail -4 gpurun_out/${TAG}_sanitizer_racecheck.log
