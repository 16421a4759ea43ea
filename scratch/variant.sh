#!/bin/bash
# builds a variant of the library with extra nvcc defines: scratch/variant.sh NAME -DFOO=1 ...  -> scratch/lib_NAME.so
set -e
NAME=$1; shift
cd "$(dirname "$0")/../ako_b200/csrc"
mkdir -p build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden \
  -I../../include -I. "$@" -c ako_device.cu -o build/ako_device_$NAME.o
[ -f build/ako_host.o ] || make build/ako_host.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scratch/lib_$NAME.so build/ako_device_$NAME.o build/ako_host.o -lpthread -lm
rm -f build/ako_device_$NAME.o
echo scratch/lib_$NAME.so
