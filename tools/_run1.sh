export C4_FZ_TIMEOUT_S=40
echo "== default nb9 g1"; timeout 100 python tools/fused_prof.py 4096 4096 --warm 2>&1 | tail -4
for v in nb9g2 nb16g2b; do echo "== $v"; C4_LIB=connect4_b200/lib/variants/libc4b200_$v.so timeout 100 python tools/fused_prof.py 4096 4096 --warm 2>&1 | tail -4; done
C4_FZ_DEBUG=1 C4_LIB=connect4_b200/lib/variants/libc4b200_nb9g2.so timeout 100 python tools/fused_prof.py 4096 4096 2>&1 | tail -2
