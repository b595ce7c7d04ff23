#!/bin/bash
# Short version of final_r01.sh: tests, 1M-block traffic of the compress kernel, both bench arms, launch list.
ulimit -c 0
O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -W ignore::DeprecationWarning > $O/gpu_tests.log 2>&1; tail -2 $O/gpu_tests.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none -k regex:lz4_compress_fast3 -s 1 -c 1 --csv --log-file $O/compress_1m_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
python tools/traffic_from_ncu.py $O/compress_1m_metrics.csv profiles/compress_traffic.json 1048576 13 "lz4_compress_fast3_kernel<13,dense>" > /dev/null && cp profiles/compress_traffic.json $O/
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err; tail -c 400 $O/bench_full.json
timeout 900 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference.err; tail -c 200 $O/bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/bench_under_ncu.log 2>&1
python - <<'PY'
import json; j=json.load(open('gpurun_out/final/bench_full.json')); print({k:j[k] for k in ('value','compress_gibs','decompress_gibs','ratio')}, j['e2e']['value'], j['cpu_baseline']['value'])
PY
