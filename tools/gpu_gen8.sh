T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
$T tools/generate.py --games 1200 --net example_config 2>&1 | tail -1
$T tools/generate.py --games 32768 --net example_config 2>&1 | tail -1
