mkdir -p gpurun_out
export PYFASST_ESTEP_KERNEL=${KERN:-ws} PYFASST_ESTEP_WSVEC=${WSVEC:-2}
CMD="python scripts/profile_driver.py --iters 2"
timeout 600 $CMD > gpurun_out/driver_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/driver_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:estep_stereo -s 1 -c 1 -f -o gpurun_out/prof_estep_${TAG:-ws} $CMD > gpurun_out/ncu_estep.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
