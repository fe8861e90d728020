export C4_FZ_TIMEOUT_S=60
timeout 200 python tools/fused_check.py --quick 2>&1 | tail -3
for n in 1024 2048 4096; do for e in fused lockstep; do C4_ENGINE=$e timeout 100 python tools/fused_prof.py $n $n 2>&1 | tail -1 | cut -c1-130; done; done
C4_ENGINE=lockstep timeout 100 python tools/fused_prof.py 8192 8192 2>&1 | tail -1 | cut -c1-130
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
