#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py tests/test_lead_sep_gpu.py tests/test_wf0_gpu.py -m gpu -q -x --timeout=300 -k "stft or istft or api or separat or wf0 or model" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stft.json 2> gpurun_out/bench_stft.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_stft.json').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v*1e3,2) for k,v in d['e2e']['stages'].items()})
PY
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:stft_kernel -c 3 --csv $CMD 2>/dev/null | grep stft_kernel | awk -F'","' '{print $5, $NF}' | cut -c1-60,200-
