#!/bin/bash
# GPU box: FP64 micro-benchmarks + ncu --set full of the two E-step kernels
mkdir -p gpurun_out
./scripts/micro/fp64_lat > gpurun_out/micro_fp64_lat.txt 2>&1
./scripts/micro/dmma > gpurun_out/micro_dmma.txt 2>&1
cat gpurun_out/micro_fp64_lat.txt gpurun_out/micro_dmma.txt
CMD="python scripts/time_estep.py --reps 2"
$CMD > gpurun_out/plain1.log 2>&1 || { echo plain failed; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:estep_stereo_kernel -s 2 -c 1 -f -o gpurun_out/r02_prof_estep_stereo_kernel $CMD > gpurun_out/ncu1.log 2>&1; echo "ncu stereo $?"
CMD="python scripts/time_estep.py --reps 2 --I 4 --conv --rank 4"
$CMD > gpurun_out/plain2.log 2>&1 || { echo plain failed; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:estep_multi_kernel -s 2 -c 1 -f -o gpurun_out/r02_prof_estep_multi_kernel $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu multi $?"
ls -la gpurun_out/*.ncu-rep
