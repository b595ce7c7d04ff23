#!/bin/bash
# Round 2: GPU parity suite + bench on the build with the wide compressor (algo 5) as the default.
ulimit -c 0
O=gpurun_out/r02h; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x -W ignore::DeprecationWarning > $O/gpu_tests.log 2>&1; tail -5 $O/gpu_tests.log | cut -c1-300
timeout 900 python bench.py > $O/bench_full.json 2> $O/bench_full.err; tail -c 1500 $O/bench_full.json; tail -3 $O/bench_full.err
