#!/bin/bash
# GPU box: the evidence set of round 2 -- GPU tests, smoke, bench (both arms, f64, configs[0], configs[2]),
# launch list, ncu --set full of the kernels that changed (each after its command exited 0 without ncu).
TAG=${TAG:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref exit $?"
python bench.py --steps 20 --warmup 3 --dtype f64 --no-cpu-baseline > gpurun_out/${TAG}_bench_f64.json 2>> gpurun_out/${TAG}_bench.err; echo "f64 exit $?"
python bench.py --steps 50 --warmup 3 --workload tamy > gpurun_out/${TAG}_bench_tamy.json 2>> gpurun_out/${TAG}_bench.err; echo "tamy exit $?"
python bench.py --steps 5 --warmup 3 --workload simm > gpurun_out/${TAG}_bench_simm.json 2>> gpurun_out/${TAG}_bench.err; echo "simm exit $?"
python bench.py --steps 20 --warmup 3 --model conv --no-cpu-baseline > gpurun_out/${TAG}_bench_conv.json 2>> gpurun_out/${TAG}_bench.err; echo "conv exit $?"
python - <<PY
import json
for name in ("bench", "bench_f64", "bench_tamy", "bench_conv"):
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json' % name).read().strip().splitlines()[-1])
        print(name, 'value %.4e e2e %.4e ms/step %.4f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']), 'roofline %.3f' % d['roofline']['frac'], d['phases_ms'])
    except Exception as e:
        print(name, 'failed', e)
for name in ("bench_reference", "bench_simm"):
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json' % name).read().strip().splitlines()[-1]); print(name, '%.4e' % d['value'], d.get('ms_per_step'))
    except Exception as e:
        print(name, 'failed', e)
PY
tail -3 gpurun_out/${TAG}_bench.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "ncu launch-list exit $?"
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1; head -12 gpurun_out/${TAG}_launches_summary.txt
CMD2="python scripts/profile_driver.py --iters 4"
$CMD2 > gpurun_out/${TAG}_driver_plain.log 2>&1 || { echo "driver failed"; exit 1; }
for kern in spec_power_tc_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$kern -s 6 -c 1 \
      -f -o gpurun_out/${TAG}_prof_$kern $CMD2 > gpurun_out/${TAG}_ncu_$kern.log 2>&1
  echo "ncu $kern exit $?"
done
CMD3="python scripts/bench_separation.py 600"
$CMD3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:^stft_kernel" -s 0 -c 1 \
      -f -o gpurun_out/${TAG}_prof_stft_kernel $CMD3 > gpurun_out/${TAG}_ncu_stft.log 2>&1; echo "ncu stft exit $?"
ls gpurun_out/${TAG}_*.ncu-rep
