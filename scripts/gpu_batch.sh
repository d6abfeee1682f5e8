mkdir -p gpurun_out
for mode in ${MODES:---sequential --sequential-graph --no-graph graph}; do
  m=$mode; [ "$mode" = "graph" ] && m=""
  python scripts/bench_batch.py --clips 32 --iters 100 $m > gpurun_out/batch_${mode}.json 2> gpurun_out/batch_${mode}.err; echo "exit $? ($mode)"
  python -c "
import json
d=json.loads(open('gpurun_out/batch_${mode}.json').read().strip().splitlines()[-1])
print(d['config']['mode'], 'value %.3e wall %.2f' % (d['value'], d['wall_s']), d['stages_s'])"
done
