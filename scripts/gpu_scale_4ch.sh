#!/bin/bash
# configs[3]: 1-hour 4-channel mixture, MultiChanNMFConv 4 sources x K=32, rank 4 -- the same
# mixture on 1 GPU (8 blocks of 450 s) and sharded over all GPUs of the box (strong scaling pair).
mkdir -p gpurun_out
NG=${NG:-8}
COMMON="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --model conv --channels 4 --rank 4 --duration-s 450"
if [ "$NG" = "1" ]; then
  timeout 900 python bench.py $COMMON --blocks 8 > gpurun_out/r02_4ch_1h_n1.json 2> gpurun_out/r02_4ch_1h_n1.err
  tail -c 600 gpurun_out/r02_4ch_1h_n1.json; tail -3 gpurun_out/r02_4ch_1h_n1.err
else
  for shard in ${SHARDS:-time freq}; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $NG $COMMON --blocks $((8 / NG)) --shard $shard > gpurun_out/r02_4ch_1h_n${NG}_$shard.json 2> gpurun_out/r02_4ch_1h_n${NG}_$shard.err
    tail -c 700 gpurun_out/r02_4ch_1h_n${NG}_$shard.json; tail -3 gpurun_out/r02_4ch_1h_n${NG}_$shard.err
  done
fi
