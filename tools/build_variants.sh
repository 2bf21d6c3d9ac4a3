#!/bin/bash
# Build librvcr.so variants that differ in compile-time knobs of one source file (A/B runs on the GPU box copy one of them over
# codename-rvc-fork-3_b200/librvcr.so).  usage: tools/build_variants.sh <source.cu> <name>:"<-D flags>" ...
set -eu
PKG=codename-rvc-fork-3_b200
SRC=$1; shift
python $PKG/build.py > /dev/null
mkdir -p variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -Wno-deprecated-declarations -diag-suppress 20012"
base=$(basename $SRC .cu)
for v in "$@"; do
  name=${v%%:*}; defs=${v#*:}
  nvcc $FLAGS $defs -I include -c $PKG/csrc/$SRC -o variants/${base}_$name.o
  objs=$(ls $PKG/_obj/*.o | grep -v "/$base.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/librvcr_$name.so $objs variants/${base}_$name.o -cudart static -Xlinker --no-undefined -lpthread -ldl -lrt
  echo "built variants/librvcr_$name.so ($defs)"
done
