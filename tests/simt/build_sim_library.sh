#!/bin/bash
# Build the WHOLE library for the CPU emulator: tests/simt/_build/libb200lz4_sim.so (test infrastructure only; see
# sim_launchers.cpp).  The product library is built by lz4-java_b200/csrc/Makefile with nvcc and is not affected.
set -e
cd "$(dirname "$0")/../.."
mkdir -p tests/simt/_build
CXX="g++ -O1 -std=c++17 -fPIC -Wno-unknown-pragmas -Wno-attributes -DB200_HOST_SIM -Itests/simt -Ilz4-java_b200/csrc"
$CXX -c tests/simt/sim_launchers.cpp -o tests/simt/_build/sim_launchers.o
for f in capi frame containers; do $CXX -x c++ -c lz4-java_b200/csrc/$f.cu -o tests/simt/_build/$f.o; done
g++ -shared -o tests/simt/_build/libb200lz4_sim.so tests/simt/_build/sim_launchers.o tests/simt/_build/capi.o tests/simt/_build/frame.o tests/simt/_build/containers.o -lpthread
echo built tests/simt/_build/libb200lz4_sim.so
