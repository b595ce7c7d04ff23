ulimit -c 0
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/p12_$i.log 2>&1; echo "full $i rc=$?"; tail -1 gpurun_out/p12_$i.log; done
VARIANTS=13:0 timeout 600 python tools/probe.py 2>&1 | tail -9
