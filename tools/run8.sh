which gdb valgrind 2>&1 | head -2
ulimit -c 0
for i in 1 2 3 4 5 6 7 8; do
  MALLOC_CHECK_=3 MALLOC_PERTURB_=165 timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning -k "decompress" > gpurun_out/p8_$i.log 2>&1
  echo "iter $i rc=$?"; tail -2 gpurun_out/p8_$i.log | cut -c1-200
done
grep -l -E "Fatal|malloc|free\(\)|corrupt" gpurun_out/p8_*.log | head
