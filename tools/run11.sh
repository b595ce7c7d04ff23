ulimit -c 0
for i in 1 2 3 4 5 6; do
  MALLOC_CHECK_=3 MALLOC_PERTURB_=165 timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning -k "decompress" > gpurun_out/p11_$i.log 2>&1
  echo "iter $i rc=$?"
done
for i in 1 2 3; do timeout 300 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/p11f_$i.log 2>&1; echo "full $i rc=$?"; tail -1 gpurun_out/p11f_$i.log; done
