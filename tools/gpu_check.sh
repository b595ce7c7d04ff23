#!/bin/bash
# One GPU-box visit during development: the GPU parity tests, then the device-resident throughput probe.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_check.sh'
ulimit -c 0
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/gpu_tests.log 2>&1
tail -3 gpurun_out/gpu_tests.log | cut -c1-300
timeout 600 python tools/probe.py 2>&1 | tail -12
