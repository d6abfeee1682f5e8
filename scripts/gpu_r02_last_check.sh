#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 2>&1 | tail -2 | tee gpurun_out/r02c_pytest_gpu.log
PYFASST_TW_FUSED=0 timeout 300 python -m pytest tests/test_tamy_gpu.py -m gpu -q -x --timeout=200 2>&1 | tail -1
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02c_bench.json').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v,3) for k,v in d['phases_ms'].items()}, round(d['roofline']['frac'],3), {k: round(v*1e3,2) for k,v in d['e2e']['stages'].items()}, d['clocks'])
PY
