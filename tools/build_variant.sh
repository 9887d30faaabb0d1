#!/bin/bash
# Development aid: builds libsift_gpu.so with extra nvcc defines into tools/variants/<name>.so (load it with SIFT_GPU_LIB=...).
#   tools/build_variant.sh nstg2 -DSL_NSTG_SMALL=2
set -e
name=$1; shift
cd "$(dirname "$0")/../sift_b200/csrc"
mkdir -p ../../tools/variants _build/v_$name
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in pyramid blur_slide extrema eliminate orient_desc sift_gpu; do
  if [ $f = blur_slide ] || [ ! -f _build/$f.o ]; then
    /usr/local/cuda/bin/nvcc $ARCH -std=c++17 -O3 -lineinfo -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC "$@" -c $f.cu -o _build/v_$name/$f.o
  else cp _build/$f.o _build/v_$name/$f.o; fi
done
/usr/local/cuda/bin/nvcc $ARCH -shared -ccbin /usr/bin/g++ -o ../../tools/variants/$name.so _build/v_$name/*.o -lpthread
echo built tools/variants/$name.so
