#!/bin/bash
# (historical: the PYFASST_GEMM_PF switch of this experiment existed at commit 39e18aa; the prefetch depth is now fixed per kernel)
# ncu --set full of gemm_tf32x3_kernel on three SIMM shapes (tensor-bound, padded-M, skinny split-K)
mkdir -p gpurun_out
export PYFASST_GEMM_PF=${PF:-2}
for sh in C_f0 C_hm D_q; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 1 -f \
    -o gpurun_out/r02_prof_gemm_$sh python scripts/micro/gemm_shapes.py --only $sh --reps 1 > gpurun_out/r02_ncu_gemm_$sh.log 2>&1
  echo "ncu $sh exit $?"; tail -2 gpurun_out/r02_ncu_gemm_$sh.log
done
ls -la gpurun_out/r02_prof_gemm_*.ncu-rep
