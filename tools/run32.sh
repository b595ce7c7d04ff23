ulimit -c 0
set -x
timeout 900 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/p32.log 2>&1; tail -3 gpurun_out/p32.log | cut -c1-200
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print({k:d[k] for k in ('value','ms_per_step','compress_gibs','decompress_gibs','ratio','e2e','gpu_launches','clocks')}); print(d['roofline']); print(d['cpu_baseline'])"
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-700 gpurun_out/bench_ref.json
# DRAM traffic of one 1M-block compress launch of the default kernel
timeout 2400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:lz4_compress_fast3 -s 1 -c 1 --csv --log-file gpurun_out/compress_1m.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu3.log 2>&1
cat gpurun_out/compress_1m.csv | tail -9
