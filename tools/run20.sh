ulimit -c 0
set -x
timeout 600 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/p20.log 2>&1; tail -3 gpurun_out/p20.log | cut -c1-200
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print({k:d[k] for k in ('value','ms_per_step','compress_gibs','decompress_gibs','ratio','roofline','e2e','cpu_baseline','gpu_launches','clocks')})"
# launch list of the same command (cold-cache, serialised: shares only)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
# full capture of one compress + one decompress launch of the same workload (1M blocks)
timeout 2400 ncu --set full --clock-control none --import-source on -k regex:lz4_ -s 2 -c 2 -o gpurun_out/prof_bench_r01 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out | tail -8
