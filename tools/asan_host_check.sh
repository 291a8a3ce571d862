#!/bin/bash
# Host-side setup code under AddressSanitizer + UBSan (no GPU needed): builds the three host sources with g++ into
# /tmp/hmg_asan and runs tools/asan_host_check.py against them.  Last run (round 2, final code): clean.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=/tmp/hmg_asan
mkdir -p $OUT
FLAGS="-std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -I$ROOT/include -I$ROOT/homogenization.jl_b200/csrc -I/usr/local/cuda/include"
for f in reference topology introspect; do g++ $FLAGS -x c++ -c $ROOT/homogenization.jl_b200/csrc/$f.cpp -o $OUT/$f.o; done
cat > $OUT/stub.cpp <<'EOC'
// the launch-shape chooser lives in kernels.cu (CUDA); the host emulation only reports its numbers
#include "kernels.cuh"
namespace hmg {
ApplyConfig make_apply_config(int, int, int, int, bool, bool) { ApplyConfig c{}; c.ring_rows = 1; return c; }
}
EOC
g++ $FLAGS -c $OUT/stub.cpp -o $OUT/stub.o
g++ -shared -fsanitize=address,undefined -o $OUT/libhmg_host_asan.so $OUT/reference.o $OUT/topology.o $OUT/introspect.o $OUT/stub.o
LD_PRELOAD=$(g++ -print-file-name=libasan.so):$(g++ -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 \
    python $ROOT/tools/asan_host_check.py $OUT/libhmg_host_asan.so
