#!/bin/bash
# build a kernel variant of libc4b200.so next to the product library: tools/build_variant.sh NAME -DFLAG...
# (selected at run time with C4_LIB=connect4_b200/lib/variants/libc4b200_NAME.so)
set -e
name=$1; shift
out=connect4_b200/lib/variants
mkdir -p $out/$name
for f in c4_board c4_search c4_net c4_fused c4_split; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 "$@" \
     -c connect4_b200/csrc/$f.cu -o $out/$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libc4b200_$name.so $out/$name/*.o
rm -r $out/$name
echo $out/libc4b200_$name.so
