#!/bin/bash
# A/B of the split-K plan of pf_gemm_tf32x3_splitk on the SIMM shapes + the pinned-parameter e2e
mkdir -p gpurun_out
for s in 296 297 148 592; do
  echo "== PYFASST_GEMM_SPLIT_SLOTS=$s"
  PYFASST_GEMM_SPLIT_SLOTS=$s timeout 120 python scripts/micro/gemm_shapes.py $( [ $s != 296 ] && echo --only split-K ) 2>&1 | tail -9
done | tee gpurun_out/gemm_shapes_ab.txt
for p in 1 0; do
  PYFASST_PINNED_PARAMS=$p timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pinned$p.json 2> gpurun_out/bench_pinned$p.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pinned$p.json').read().strip().splitlines()[-1])
print('pinned=$p value %.4e e2e %.4e step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v*1e3,2) for k,v in d['e2e']['stages'].items()})
PY
done
