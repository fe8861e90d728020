export C4_FZ_TIMEOUT_S=60
timeout 200 python tools/fused_check.py --quick 2>&1 | tail -3
for n in 256 1024 2048 4096; do C4_ENGINE=fused timeout 100 python tools/fused_prof.py $n $n 2>&1 | tail -1 | cut -c1-120; done
C4_ENGINE=fused timeout 100 python tools/fused_prof.py 4096 4096 --warm 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_selfplay.py -x -q 2>&1 | tail -3
