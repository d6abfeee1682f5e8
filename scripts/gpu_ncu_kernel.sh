# usage: KREGEX=<kernel regex> TAG=<name> [ENVV="A=1 B=2"] bash scripts/gpu_ncu_kernel.sh
mkdir -p gpurun_out
export $ENVV
CMD="python scripts/profile_driver.py --iters 2"
timeout 600 $CMD > gpurun_out/driver_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/driver_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 1 -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_$TAG.ncu-rep
