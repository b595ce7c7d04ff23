#!/bin/bash
# Round-end measurement pass on the GPU box: tests, 1M-block DRAM traffic of the two hot kernels (-> the JSON bench.py
# reads), both bench arms, launch list, full ncu captures (8192-block batches), secondary configs.
# Outputs under gpurun_out/final/.
ulimit -c 0
O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -W ignore::DeprecationWarning > $O/gpu_tests.log 2>&1; tail -2 $O/gpu_tests.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none -k regex:lz4_compress_fast3 -s 1 -c 1 --csv --log-file $O/compress_1m_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:lz4_decompress_fast -s 1 -c 1 --csv --log-file $O/decompress_1m_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
python tools/traffic_from_ncu.py $O/compress_1m_metrics.csv profiles/compress_traffic.json 1048576 13 "lz4_compress_fast3_kernel<13,dense>" > /dev/null && cp profiles/compress_traffic.json $O/
python tools/traffic_from_ncu.py $O/decompress_1m_metrics.csv $O/decompress_traffic.json 1048576 13 "lz4_decompress_fast_kernel<4,batched>" | cut -c1-300
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err; tail -c 700 $O/bench_full.json
timeout 900 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference.err; tail -c 300 $O/bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/bench_under_ncu.log 2>&1
export NBLK=8192 VARIANTS=13:0:3:0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_compress_fast3 -s 2 -c 1 -o $O/prof_compress python tools/probe.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_decompress_fast -s 2 -c 1 -o $O/prof_decompress_fast python tools/probe.py > /dev/null 2>&1
unset NBLK VARIANTS
timeout 1500 python tools/bench_configs.py > $O/secondary_configs.log 2>&1; tail -5 $O/secondary_configs.log
ls -la $O
