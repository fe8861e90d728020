#!/bin/bash
# one GPU call: parity tests, default bench line, ncu launch list of the same command, one full capture of the tree pass
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_default.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --preroll 8000 --passes 100 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_memo.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 16640 -c 200 --csv --log-file gpurun_out/r01_launches_v3.csv $CMD > gpurun_out/ncu_l.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_advance -s 8350 -c 2 -o gpurun_out/r01_advance_v3 -f $CMD > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
