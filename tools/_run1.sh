export C4_FZ_TIMEOUT_S=40
timeout 100 python tools/fused_check.py --quick 2>&1 | tail -3
echo "== default (tree high)"; timeout 100 python tools/fused_prof.py 4096 4096 --warm 2>&1 | tail -4
echo "== treelow"; C4_LIB=connect4_b200/lib/variants/libc4b200_treelow.so timeout 100 python tools/fused_prof.py 4096 4096 --warm 2>&1 | tail -4
C4_FZ_DEBUG=1 timeout 100 python tools/fused_prof.py 4096 4096 2>&1 | tail -2
for n in 256 1024 2048 8192; do for e in fused lockstep; do C4_ENGINE=$e timeout 100 python tools/fused_prof.py $n $n 2>&1 | tail -1; done; done
