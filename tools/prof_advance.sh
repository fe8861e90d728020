CMD="python bench.py --steps 1 --warmup 3 --preroll 8000 --passes 100 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_p.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:k_advance -s 8350 -c 1 -o gpurun_out/r01_advance_v4 -f $CMD > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
