#!/bin/bash
mkdir -p gpurun_out
sh=${SH:-D_q}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 1 -f \
  -o gpurun_out/r02_prof_gemm_ws_$sh python scripts/micro/gemm_shapes.py --only "$sh = " --reps 1 > gpurun_out/r02_ncu_gemm_ws_$sh.log 2>&1
echo "ncu $sh exit $?"; tail -2 gpurun_out/r02_ncu_gemm_ws_$sh.log
