ulimit -c 0
timeout 900 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/p24.log 2>&1; tail -15 gpurun_out/p24.log | cut -c1-250
