#!/bin/bash
# ncu evidence for the split engine (one launch, k_sp_one): launch list of the bench command, then one --set full capture of a
# shortened cold generation (256 games on 4,096 slots, 256 MB memo so that ncu's save / restore between replays stays small)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras"
$B > gpurun_out/r02_split_plain_bench.log 2>&1 || { tail -5 gpurun_out/r02_split_plain_bench.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_split_launches_bench.csv $B > gpurun_out/ncu_l.log 2>&1
tail -1 gpurun_out/ncu_l.log | cut -c1-300
export C4_MEMO_LOG2=22
C="python tools/fused_prof.py 256 4096"
$C > gpurun_out/r02_split_plain.log 2>&1 || { tail -5 gpurun_out/r02_split_plain.log; exit 1; }
cat gpurun_out/r02_split_plain.log | tail -1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_sp_one -c 1 -o gpurun_out/r02_split_one -f $C > gpurun_out/ncu_s.log 2>&1
tail -3 gpurun_out/ncu_s.log
