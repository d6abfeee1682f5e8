#!/bin/bash
# (historical: the PYFASST_GEMM_PF switch of this experiment existed at commit 39e18aa; the prefetch depth is now fixed per kernel)
# prefetch depth A/B of gemm_tf32x3_kernel on the SIMM shapes + its parity tests at each depth
mkdir -p gpurun_out
for pf in 2 4 1; do
  echo "== PYFASST_GEMM_PF=$pf"
  PYFASST_GEMM_PF=$pf timeout 300 python -m pytest tests/test_tc_gpu.py tests/test_simm_gpu.py -m gpu -q -x --timeout=200 -k "gemm" 2>&1 | tail -2
  PYFASST_GEMM_PF=$pf timeout 120 python scripts/micro/gemm_shapes.py 2>&1 | tail -7
done | tee gpurun_out/gemm_shapes_pf.txt
