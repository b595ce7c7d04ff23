ulimit -c 0
timeout 900 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning -k "compress or roundtrip" > gpurun_out/p23.log 2>&1; tail -3 gpurun_out/p23.log | cut -c1-250
VARIANTS=13:0:3,12:0:3,13:0:2 timeout 600 python tools/probe.py 2>&1 | head -4
