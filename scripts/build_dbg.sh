#!/bin/bash
# Instrumented twin of the library: the fused IK kernel with its per-phase cycle counters (-DSMPLPP_IK2_DBG).
# Use with SMPLPP_B200_LIB=smplpp_b200/libsmplpp_b200_dbg.so
set -e
cd "$(dirname "$0")/.."
python -c "from smplpp_b200 import build; build.build()"
mkdir -p smplpp_b200/build/dbg
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC \
  -DSMPLPP_IK2_DBG -c smplpp_b200/csrc/ik2.cu -o smplpp_b200/build/dbg/ik2.o
objs=$(ls smplpp_b200/build/*.o | grep -v "/ik2.o")
/usr/local/cuda/bin/nvcc -shared -o smplpp_b200/libsmplpp_b200_dbg.so $objs smplpp_b200/build/dbg/ik2.o -lcudart
ls -la smplpp_b200/libsmplpp_b200_dbg.so
