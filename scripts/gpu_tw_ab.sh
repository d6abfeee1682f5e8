timeout -s KILL 150 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q -x -k "tw_contract or engine" 2>&1 | tail -1
PYFASST_B200_LIB=pyfasst_b200/libpyfasst_b200_wpt.so timeout -s KILL 150 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q -x -k "tw_contract or engine" 2>&1 | tail -1
for lib in "" pyfasst_b200/libpyfasst_b200_wpt.so; do
  [ -n "$lib" ] && export PYFASST_B200_LIB=$lib
  timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=$lib %.4e %.3f ms' % (d['value'], d['ms_per_step']), d['phases_ms'], d['loglik_last'])"
done
