mkdir -p gpurun_out/det
for i in 1 2 3 4 5 6 7 8 9 10; do python tools/generate.py --games 300 --sims 200 --net default --slots 300 --dump gpurun_out/det/r$i.npy 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['positions'], d['records_sha256_16'])"; done
