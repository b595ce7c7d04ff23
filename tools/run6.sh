set -x
cat /sys/fs/cgroup/cpu.max; nproc
timeout 600 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/pytest6.log 2>&1; tail -3 gpurun_out/pytest6.log
VARIANTS=13:0,12:0 timeout 600 python tools/probe.py 2>&1 | tail -12
timeout 900 python bench.py --blocks 262144 --steps 3 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; python -c "
import json; d=json.load(open('gpurun_out/bench6.json')); print({k:d[k] for k in ('value','compress_gibs','decompress_gibs','ratio','e2e','cpu_baseline')})"
tail -3 gpurun_out/bench6.err
