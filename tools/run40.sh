ulimit -c 0
timeout 900 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/p40.log 2>&1; tail -3 gpurun_out/p40.log | cut -c1-300
VARIANTS=13:0:3:0 timeout 600 python tools/probe.py 2>&1 | grep decompress
for v in d16 d10 d8; do echo "== $v"; B200LZ4_SO=variants/libb200lz4_$v.so VARIANTS=13:0:3:0 timeout 600 python tools/probe.py 2>&1 | grep decompress; done
