ulimit -c 0
timeout 900 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning -k "hc or contract" > gpurun_out/p26.log 2>&1; tail -5 gpurun_out/p26.log | cut -c1-250
timeout 900 python tools/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; tail -45 gpurun_out/configs.json; tail -5 gpurun_out/configs.err
timeout 900 python bench.py --blocks 262144 --steps 4 --warmup 3 > gpurun_out/bench25.json 2> gpurun_out/bench25.err; tail -3 gpurun_out/bench25.err
python -c "
import json; d=json.load(open('gpurun_out/bench25.json')); print({k:d[k] for k in ('value','compress_gibs','decompress_gibs','ratio','e2e')}); print(d['cpu_baseline'])"
