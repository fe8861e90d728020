python tools/net_time.py --net64 > gpurun_out/net64_plain.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:k_net_tc -s 230 -c 1 -o gpurun_out/r01_net_tc64 -f python tools/net_time.py --net64 > gpurun_out/ncu_n64.log 2>&1
tail -12 gpurun_out/net64_plain.log; tail -2 gpurun_out/ncu_n64.log
