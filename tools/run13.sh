ulimit -c 0
timeout 600 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/p13.log 2>&1; tail -15 gpurun_out/p13.log | cut -c1-220
VARIANTS=13:0,12:0 timeout 600 python tools/probe.py 2>&1 | tail -9
