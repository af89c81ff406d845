#!/bin/bash
# builds a variant of the product library with extra nvcc flags (A/B measurements only; never loaded by default):
#   tools/build_variant.sh p1 -DFL_TILE_PLANES=1   ->  fluca_b200/csrc/variants/libfluca_b200_p1.so
set -e
name=$1; shift
cd "$(dirname "$0")/../fluca_b200/csrc"
mkdir -p variants/$name
for f in geom step krylov mg tiles ibm capi comm_nccl fd; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo --extended-lambda -Xcompiler -fPIC -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -I/usr/include "$@" -c $f.cu -o variants/$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o variants/libfluca_b200_$name.so variants/$name/*.o -L/usr/lib/x86_64-linux-gnu -lnccl -lcudart
rm -rf variants/$name
ls -la variants/libfluca_b200_$name.so
