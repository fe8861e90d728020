#!/bin/bash
# round 2 evidence run: launch list of the bench command, full captures of the lock-step engine's two kernels and of the fused kernel
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras"
$CMD > gpurun_out/plain_r02.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 20000 -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
ncu --set full --import-source on --clock-control none -k regex:k_advance -s 10000 -c 2 -o gpurun_out/r02_advance -f $CMD > gpurun_out/ncu_a.log 2>&1
tail -2 gpurun_out/ncu_a.log
ncu --set full --import-source on --clock-control none -k regex:k_net_tc -s 10000 -c 2 -o gpurun_out/r02_net_tc -f $CMD > gpurun_out/ncu_n.log 2>&1
tail -2 gpurun_out/ncu_n.log
export C4_MEMO_LOG2=22 C4_ENGINE=fused
python tools/fused_prof.py 256 4096 > gpurun_out/plain_fused.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:k_fused -c 1 -o gpurun_out/r02_fused -f python tools/fused_prof.py 256 4096 > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/plain_fused.log; tail -2 gpurun_out/ncu_f.log
