ulimit -c 0
timeout 900 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning -k "compress_roundtrip" > gpurun_out/p30.log 2>&1; tail -12 gpurun_out/p30.log | cut -c1-300
VARIANTS=13:0:4:0,13:0:3:0,12:0:4:1 timeout 600 python tools/probe.py 2>&1 | head -4
