#!/bin/bash
# GPU box: the evidence set of a round: GPU tests, smoke, bench (both arms), launch list, ncu
# --set full captures of the dominant kernels (each after its command exited 0 without ncu).
TAG=${TAG:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e ms/step %.3f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']))
print('e2e stages', d['e2e'].get('stages')); print('phases', d['phases_ms']); print('roofline', d['roofline']); print('clocks', d['clocks'])
r=json.loads(open('gpurun_out/${TAG}_bench_reference.json').read().strip().splitlines()[-1]); print('reference arm value %.4e' % r['value'])
"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_bench_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "ncu launch-list exit $?"
CMD2="python scripts/profile_driver.py --iters 4"
$CMD2 > gpurun_out/${TAG}_driver_plain.log 2>&1 || { echo "driver failed"; exit 1; }
for kern in ${KERNELS:-estep_stereo_kernel tw_contract_fused_tc_kernel fb_contract_tc_kernel spec_power_tc_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$kern -s 2 -c 1 \
      -f -o gpurun_out/${TAG}_prof_$kern $CMD2 > gpurun_out/${TAG}_ncu_$kern.log 2>&1
  echo "ncu $kern exit $?"
done
[ "${SKIP_SIMM:-0}" = "1" ] && exit 0
# SIMM (configs[2]): bench line, launch list, and the dense contraction kernel
python scripts/bench_simm.py --steps 3 --warmup 3 > gpurun_out/${TAG}_simm_bench.json 2> gpurun_out/${TAG}_simm_bench.err; echo "simm bench exit $?"
CMD3="python scripts/bench_simm.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${TAG}_simm_launches.csv $CMD3 > gpurun_out/${TAG}_simm_ncu_launch.log 2>&1
echo "simm launch-list exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_kernel -s 20 -c 1 \
    -f -o gpurun_out/${TAG}_prof_gemm_tf32x3_kernel $CMD3 > gpurun_out/${TAG}_ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
ls gpurun_out/${TAG}_*.ncu-rep
