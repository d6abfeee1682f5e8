#!/bin/bash
# GPU box: ncu --set full captures of the kernels that had no committed capture (each after its
# command exited 0 without ncu), and compute-sanitizer memcheck / racecheck over kernel tests.
TAG=${TAG:-r02}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
CMD="python scripts/profile_driver.py --iters 3"
timeout 600 $CMD > gpurun_out/${TAG}_driver_plain.log 2>&1 || { echo "driver failed"; tail -5 gpurun_out/${TAG}_driver_plain.log; }
for kern in tw_contract_fused_tc_kernel spec_power_tc_kernel fb_contract_tc_kernel; do
  timeout 600 $NCU -k regex:$kern -s 6 -o gpurun_out/${TAG}_prof_$kern $CMD > gpurun_out/${TAG}_ncu_$kern.log 2>&1; echo "ncu $kern $?"
done
CMD="python scripts/bench_separation.py 600"
timeout 600 $CMD > gpurun_out/${TAG}_separation.json 2> gpurun_out/${TAG}_separation.err || echo "separation failed"
tail -1 gpurun_out/${TAG}_separation.json
for kern in "^stft_kernel" istft_kernel wiener_stereo_kernel; do
  timeout 900 $NCU -k regex:$kern -s 1 -o gpurun_out/${TAG}_prof_$kern $CMD > gpurun_out/${TAG}_ncu_$kern.log 2>&1; echo "ncu $kern $?"
done
CMD="python scripts/bench_viterbi.py"
timeout 600 $CMD > gpurun_out/${TAG}_viterbi.txt 2>&1 || echo "viterbi failed"
tail -4 gpurun_out/${TAG}_viterbi.txt
timeout 900 $NCU -k regex:viterbi -s 1 -o gpurun_out/${TAG}_prof_viterbi_kernel $CMD > gpurun_out/${TAG}_ncu_viterbi.log 2>&1; echo "ncu viterbi $?"
# compute-sanitizer over the small-fixture kernel tests
SAN="compute-sanitizer --error-exitcode 86 --launch-timeout 0"
timeout 1500 $SAN --tool memcheck python -m pytest tests/test_kernels_gpu.py tests/test_multichannel_gpu.py -m gpu -q -x --timeout=1400 -k "estep or wiener or contract or spec_power or mix or stft" > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"
tail -4 gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 1500 $SAN --tool racecheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=1400 -k "estep or contract or spec_power" > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"
tail -4 gpurun_out/${TAG}_sanitizer_racecheck.log
ls -la gpurun_out/${TAG}_prof_*.ncu-rep
