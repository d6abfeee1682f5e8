#!/bin/bash
# is the ~2.5 TB/s ceiling of the streaming shapes a TLB effect of the 1.65 MB row stride?
mkdir -p gpurun_out
for n in 103362 12000; do
  echo "== frames $n"
  timeout 120 python scripts/micro/gemm_shapes.py --frames $n --only D_q,C_phi,tn 2>&1 | tail -5
done | tee gpurun_out/gemm_shapes_tlb.txt
