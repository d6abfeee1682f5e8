mkdir -p gpurun_out
python scripts/bench_simm.py --steps 3 --warmup 3 > gpurun_out/simm_bench.json 2> gpurun_out/simm_bench.err; echo "exit $?"
tail -3 gpurun_out/simm_bench.err; cat gpurun_out/simm_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/simm_launches.csv python scripts/bench_simm.py --frames 103362 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/simm_ncu.log 2>&1; echo "ncu exit $?"
python scripts/launch_summary.py gpurun_out/simm_launches.csv | head -30
