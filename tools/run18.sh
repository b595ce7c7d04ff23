ulimit -c 0
timeout 900 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/p18.log 2>&1; tail -12 gpurun_out/p18.log | cut -c1-250
VARIANTS=13:0:2,12:0:2,13:0:1 timeout 600 python tools/probe.py 2>&1 | tail -11
export NBLK=8192 VARIANTS=13:0:2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_compress_fast2 -s 2 -c 1 -o gpurun_out/prof_compress_r1e python tools/probe.py > gpurun_out/ncu_c.log 2>&1
