#!/bin/bash
# Round 2: the two repaired GPU tests, then a short bench run with every leg on (reduced sizes: a debug pass of the new keys).
ulimit -c 0
O=gpurun_out/r02k; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -W ignore::DeprecationWarning -k "pinned or negative" > $O/gpu_tests.log 2>&1; tail -3 $O/gpu_tests.log | cut -c1-300
timeout 900 python bench.py --blocks 65536 --xxh-buffers 1000000 --frames 8 --hc-blocks 4096 --steps 2 --warmup 3 --e2e-blocks 8192 --cpu-blocks 2048 > $O/bench_small.json 2> $O/bench_small.err; tail -c 4000 $O/bench_small.json; tail -5 $O/bench_small.err
