#!/bin/bash
# GPU box: launch list of one source/filter GEM run + ncu --set full of its plane kernel and of the
# F0-dictionary kernel
mkdir -p gpurun_out
CMD="python scripts/bench_sourcefilter.py --iters 2 --warmup 1"
timeout 600 $CMD > gpurun_out/sf_plain.log 2>&1 || { echo plain failed; tail -3 gpurun_out/sf_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r01_sourcefilter_launches.csv $CMD > gpurun_out/sf_ncu_launch.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gem_ratio_planes_kernel -s 4 -c 1 -f -o gpurun_out/prof_gem_ratio_planes $CMD > gpurun_out/sf_ncu1.log 2>&1; echo "ncu planes exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wf0_comb_kernel -c 1 -f -o gpurun_out/prof_wf0_comb $CMD > gpurun_out/sf_ncu2.log 2>&1; echo "ncu wf0 exit $?"
