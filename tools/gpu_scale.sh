#!/bin/bash
# 8-GPU box: the bench at N = 8 and N = 4 (weak scaling, one pool per GPU), then one sharded generation at N = 8
mkdir -p gpurun_out
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n$n.json
  python -c "import json; d=json.load(open('gpurun_out/bench_n$n.json')); print($n, d['value'], d['e2e']['value'], d['clocks'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/generate.py --games 32768 --net default 2>&1 | tail -1 | tee gpurun_out/gen_default_n8.json
