#!/bin/bash
# full GPU test suite + SIMM / main bench lines after the GEMM changes of round 2 (late)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 2>&1 | tail -3 | tee gpurun_out/r02b_pytest_gpu.log
PYFASST_GEMM_TC=1 timeout 200 python -m pytest tests/test_tc_gpu.py -m gpu -q -x -k gemm 2>&1 | tail -1
timeout 120 python scripts/micro/gemm_shapes.py --only C_phi,C_hm,SM_c 2>&1 | tail -4 | tee gpurun_out/gemm_shapes_ws3.txt
timeout 300 python bench.py --workload simm --steps 5 --warmup 3 > gpurun_out/r02b_bench_simm.json 2> gpurun_out/r02b_bench_simm.err; tail -c 600 gpurun_out/r02b_bench_simm.json | head -c 400; echo
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_chk$i.json 2> gpurun_out/bench_chk$i.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_chk$i.json').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v,3) for k,v in d['phases_ms'].items()}, round(d['roofline']['frac'],3), round(d['e2e']['stages']['estim.gem_s']*1e3,2))
PY
done
