#!/bin/bash
# GPU box: ncu --set full of the stereo E-step launch of scripts/time_estep.py (ARGS, ENVV optional)
mkdir -p gpurun_out
[ -n "$ENVV" ] && export $ENVV
CMD="python scripts/time_estep.py --reps 2 $ARGS"
$CMD > gpurun_out/plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain.log; exit 1; }
tail -1 gpurun_out/plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-estep_stereo} -s 2 -c 1 -f -o gpurun_out/${TAG:-prof_estep} $CMD > gpurun_out/ncu_${TAG:-prof_estep}.log 2>&1; echo "ncu $?"
