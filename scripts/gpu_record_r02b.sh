#!/bin/bash
# GPU box: evidence set at the final code of round 2 (after the GEMM / tf32-split / pinned-parameter
# changes): smoke, bench lines, launch list, ncu of the dedicated-MMA-warp GEMM.  (The GPU test run of
# the same code: scripts/gpu_r02_final_check.sh -> gpurun_out/r02b_pytest_gpu.log.)
TAG=${TAG:-r02b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref exit $?"
timeout 300 python bench.py --steps 20 --warmup 3 --dtype f64 --no-cpu-baseline > gpurun_out/${TAG}_bench_f64.json 2>> gpurun_out/${TAG}_bench.err; echo "f64 exit $?"
timeout 300 python bench.py --steps 50 --warmup 3 --workload tamy --no-cpu-baseline > gpurun_out/${TAG}_bench_tamy.json 2>> gpurun_out/${TAG}_bench.err; echo "tamy exit $?"
timeout 300 python bench.py --steps 20 --warmup 3 --model conv --no-cpu-baseline > gpurun_out/${TAG}_bench_conv.json 2>> gpurun_out/${TAG}_bench.err; echo "conv exit $?"
python - <<PY
import json
for name in ("bench", "bench_f64", "bench_tamy", "bench_conv"):
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json' % name).read().strip().splitlines()[-1])
        print(name, 'value %.4e e2e %.4e ms/step %.4f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']), 'roofline %.3f' % d['roofline']['frac'], {k: round(v, 3) for k, v in d['phases_ms'].items()})
    except Exception as e:
        print(name, 'failed', e)
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_reference.json').read().strip().splitlines()[-1]); print('reference', '%.4e' % d['value'], d.get('ms_per_step'))
except Exception as e:
    print('reference failed', e)
PY
tail -3 gpurun_out/${TAG}_bench.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "ncu launch-list exit $?"
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1; head -8 gpurun_out/${TAG}_launches_summary.txt
for sh in C_f0 D_q; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 1 -f \
    -o gpurun_out/${TAG}_prof_gemm_ws_$sh python scripts/micro/gemm_shapes.py --only "$sh = " --reps 1 > gpurun_out/${TAG}_ncu_gemm_ws_$sh.log 2>&1
  echo "ncu $sh exit $?"
done
