mkdir -p gpurun_out
export PYFASST_ESTEP_VARIANT=${V:-3}
CMD="python scripts/profile_driver.py --iters 2"
$CMD > gpurun_out/driver_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/driver_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:estep_stereo_kernel -s 1 -c 1 -f -o gpurun_out/prof_estep_v${V:-3} $CMD > gpurun_out/ncu_estep.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
