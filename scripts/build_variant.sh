#!/bin/bash
# build_variant.sh <libname.so> "<extra nvcc flags>" file1.cu [file2.cu ...]
# Recompiles the named sources with extra flags and links them with the standard objects into rrtqx_3d_b200/<libname.so>
# (A/B experiments: RRTQX_B200_LIB=$PWD/rrtqx_3d_b200/<libname.so>).
set -e
cd "$(dirname "$0")/../rrtqx_3d_b200/csrc"
lib=$1; extra=$2; shift 2
objs=""
for f in abi tree range collision sweep sweep2d polygon dubins comm; do
  if [[ " $* " == *" $f.cu "* ]]; then
    /usr/local/cuda/bin/nvcc $extra -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
      -Xcompiler -fPIC,-fvisibility=hidden,-Wall -Xptxas -v --expt-relaxed-constexpr -c $f.cu -o /tmp/var_${lib}_$f.o 2> /tmp/var_${lib}_$f.log
    objs="$objs /tmp/var_${lib}_$f.o"
  else
    objs="$objs $f.o"
  fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../$lib $objs -ldl
echo built ../$lib
