#!/bin/bash
# Build alternative libb200lz4.so files with different compile-time knobs of the compress / decompress kernels
# (for A/B probes on the GPU box: B200LZ4_TEST_SO=variants/libb200lz4_<name>.so python tools/probe.py).
# usage: tools/build_variants.sh name:"-DFLAG=.. -DFLAG=.." ...
set -e
cd "$(dirname "$0")/../lz4-java_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  for f in lz4_compress lz4_decompress; do
    nvcc -std=c++17 -O3 -lineinfo $ARCH -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c $f.cu -o /tmp/${f}_$name.o
  done
  nvcc $ARCH -shared -o ../../variants/libb200lz4_$name.so capi.o /tmp/lz4_decompress_$name.o /tmp/lz4_compress_$name.o lz4hc_compress.o xxhash.o compact.o frame.o containers.o -Xlinker --version-script=exports.map
  echo "built variants/libb200lz4_$name.so ($flags)"
done
