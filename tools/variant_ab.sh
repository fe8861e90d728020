#!/bin/bash
# A/B of library variants (tools/build_variant.sh) in ONE call on ONE box: cold benchmark generation, constant 72 towers
# (tools/phase_ab.py --only-tables), alternating processes.  usage: variant_ab.sh name1 name2 ...   ("" = the product library)
mkdir -p gpurun_out
for i in 1 2; do
  for v in "" "$@"; do
    lib=""; [ -n "$v" ] && lib=connect4_b200/lib/variants/libc4b200_$v.so
    echo "variant=${v:-product} $(C4_LIB=$lib timeout 100 python tools/phase_ab.py --reps 1 --only-tables 2>&1 | grep '^tables 1' | awk '{printf "%s ", $3}')"
  done
done
