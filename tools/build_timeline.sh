#!/bin/sh
# Debug library for tools/timeline.py: flat_search.cu with -DBP_TIMELINE (every warp of the
# launch-per-depth kernels stamps %globaltimer), linked with the shipped objects.  Not shipped.
set -e
cd "$(dirname "$0")/.."
python -m visapi_b200.build > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC --threads 2 \
     -DBP_TIMELINE -c visapi_b200/csrc/flat_search.cu -o /tmp/flat_search_tl.o
nvcc -shared -o visapi_b200/librectpack_b200_timeline.so /tmp/flat_search_tl.o \
     visapi_b200/csrc/build/api_basic.o visapi_b200/csrc/build/puct.o visapi_b200/csrc/build/resident_k22.o \
     visapi_b200/csrc/build/resident_k24.o visapi_b200/csrc/build/resident_k42.o visapi_b200/csrc/build/resident_k44.o \
     -gencode arch=compute_100a,code=sm_100a
echo visapi_b200/librectpack_b200_timeline.so
