#!/bin/bash
# builds the TEST-ONLY museum library next to this script (never shipped, never linked by libsrw.so)
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
LIBDIR=$(cd ../../stellar-random-walk_b200 && pwd)
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ --expt-relaxed-constexpr --extended-lambda \
  -shared -o libsrw_museum.so walk_museum.cu -L"$LIBDIR" -lsrw -Xlinker -rpath -Xlinker "$LIBDIR" -cudart static
echo "$(pwd)/libsrw_museum.so"
