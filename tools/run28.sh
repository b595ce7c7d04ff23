ulimit -c 0
which compute-sanitizer
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning -k "not jni" > gpurun_out/sanitizer.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitizer.log | head -20
