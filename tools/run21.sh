ulimit -c 0
timeout 900 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/p21.log 2>&1; tail -12 gpurun_out/p21.log | cut -c1-250
timeout 600 python tools/probe.py 2>&1 | tail -13
