#!/bin/bash
# round 2, last state (shared-memory PUCT tables + adaptive tower count): GPU parity tests, the default bench line, then -- with
# `ncu` as first argument -- the ncu launch list (+ DRAM bytes) of the short bench command and one --set full capture of a slice
mkdir -p gpurun_out
if [ "$1" != "ncu" ]; then
  ( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_final_gputests.log 2>&1; tail -4 gpurun_out/r02_final_gputests.log
  ( time python bench.py ) > gpurun_out/r02_final_bench.log 2>&1; tail -4 gpurun_out/r02_final_bench.log | cut -c1-1500
  exit 0
fi
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras"
$B > gpurun_out/r02_adapt_plain_bench.log 2>&1 || { tail -5 gpurun_out/r02_adapt_plain_bench.log; exit 1; }
tail -1 gpurun_out/r02_adapt_plain_bench.log | cut -c1-200
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02_split_adapt_launches.csv $B > gpurun_out/ncu_l.log 2>&1
tail -1 gpurun_out/ncu_l.log | cut -c1-300
export C4_MEMO_LOG2=22
C="python tools/fused_prof.py 1024 4096"
$C > gpurun_out/r02_adapt_plain.log 2>&1 || { tail -5 gpurun_out/r02_adapt_plain.log; exit 1; }
tail -1 gpurun_out/r02_adapt_plain.log
# the 6th slice of the generation: the middle phase (towers saturated, ~80 of them)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_sp_one -s 5 -c 1 -o gpurun_out/r02_split_adapt -f $C > gpurun_out/ncu_s.log 2>&1
tail -3 gpurun_out/ncu_s.log
