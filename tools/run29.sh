ulimit -c 0
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning -k "compress_roundtrip or hc or xxhash_uniform" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/racecheck.log | head -20
