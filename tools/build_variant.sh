#!/bin/bash
# build a library variant with extra nvcc flags into variants/libpm3d_<name>.so (git-ignored; travels with gpurun):
#   tools/build_var_$nameiant.sh name "-DCOMPOSE_FLOOR_NP=1 ..."
set -e
cd "$(dirname "$0")/../pixelmek-3d_b200/csrc"
name=$1; shift
rm -rf build_var_$name && mkdir -p build_var_$name
for f in pm3d_api.cu front_kernel.cu compose_kernel.cu los_kernel.cu present_kernel.cu clutter_kernel.cu sprite_list_kernel.cu selfcheck.cu comm.cu path_kernel.cu wall_lines.cpp; do
  if [ -n "$ONLY" ] && [[ " $ONLY " != *" $f "* ]] && [ -f build/$f.o ]; then cp build/$f.o build_var_$name/$f.o; continue; fi
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC,-Wall -I../../include -I. $@ -c -o build_var_$name/$f.o $f &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/libpm3d_$name.so build_var_$name/*.o -lcudart -ldl
rm -rf build_var_$name
ls -la ../../variants/libpm3d_$name.so
