ulimit -c 0
timeout 900 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/p43.log 2>&1; tail -4 gpurun_out/p43.log | cut -c1-400
VARIANTS=13:0:3:0 timeout 600 python tools/probe.py 2>&1 | grep -E "decompress"
ONLY=config3 timeout 900 python tools/bench_configs.py 2>&1 | tail -12
NBLK=512 BS=4194304 VARIANTS=12:0:3:0 timeout 600 python tools/probe.py 2>&1 | grep -E "decompress"
NBLK=2048 VARIANTS=13:0:3:0 timeout 600 python tools/probe.py 2>&1 | grep -E "decompress"
