#!/bin/bash
# warp-specialised GEMM (dedicated MMA warp) against the barrier-per-chunk kernel on the SIMM shapes
mkdir -p gpurun_out
for ws in 1 0; do
  echo "== PYFASST_GEMM_WS=$ws"
  PYFASST_GEMM_WS=$ws timeout 200 python -m pytest tests/test_tc_gpu.py tests/test_simm_gpu.py -m gpu -q -x --timeout=100 -k "gemm" 2>&1 | tail -3
  PYFASST_GEMM_WS=$ws timeout 120 python scripts/micro/gemm_shapes.py 2>&1 | tail -7
done | tee gpurun_out/gemm_shapes_ws.txt
