#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (run on the GPU box; outputs under gpurun_out/).
ulimit -c 0
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -W ignore::DeprecationWarning -k "not jni" > gpurun_out/sanitizer_memcheck.txt 2>&1
tail -4 gpurun_out/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -W ignore::DeprecationWarning -k "decompress or xxhash_long or (compress_roundtrip and u16)" > gpurun_out/sanitizer_racecheck.txt 2>&1
grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_racecheck.txt | tail -3
grep -E "hazard detected" -A3 gpurun_out/sanitizer_racecheck.txt | grep -E "at .*\.cu" | sed 's/.* in //' | sort | uniq -c | sort -rn | head -12
