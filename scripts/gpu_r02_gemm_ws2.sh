#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_tc_gpu.py tests/test_simm_gpu.py tests/test_kernels_gpu.py -m gpu -q -x --timeout=100 2>&1 | tail -3
timeout 120 python scripts/micro/gemm_shapes.py 2>&1 | tail -9 | tee gpurun_out/gemm_shapes_ws2.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_split.json').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['phases_ms'], d['roofline']['frac'])
PY
