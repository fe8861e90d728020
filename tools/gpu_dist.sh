#!/bin/bash
# 2-GPU box: world-size independence of a sharded generation over real NCCL, then the bench at N=2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
python tools/generate.py --games 1200 --net default 2>&1 | tail -1 | tee gpurun_out/gen_default_n1.json
$T tools/generate.py --games 1200 --net default 2>&1 | tail -1 | tee gpurun_out/gen_default_n2.json
python tools/generate.py --games 1200 --net example_config --out gpurun_out/gen_out 2>&1 | tail -1 | tee gpurun_out/gen_big_n1.json
$T tools/generate.py --games 1200 --net example_config 2>&1 | tail -1 | tee gpurun_out/gen_big_n2.json
rm -rf gpurun_out/gen_out
if [ "$1" = "bench" ]; then $T bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n2.json; fi
