#!/bin/bash
# Round 2, final visit to a 1-GPU box: everything profiles/ quotes for the round's last build.
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/r02_final.sh'
ulimit -c 0
O=gpurun_out/r02final; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -W ignore::DeprecationWarning > $O/gpu_tests.log 2>&1; tail -3 $O/gpu_tests.log | cut -c1-300
timeout 600 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference.err; tail -c 300 $O/bench_reference_arm.json
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err; tail -c 600 $O/bench_full.json
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > $O/bench_under_ncu.log 2>&1
# DRAM traffic + counters of the 1 M-block launches
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --metrics $M --clock-control none -k regex:lz4_compress_wide -s 1 -c 1 --csv --log-file $O/compress_1m_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary > /dev/null 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:lz4_decompress_fast -s 1 -c 1 --csv --log-file $O/decompress_1m_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary > /dev/null 2>&1
# full captures (8192 blocks) of the two kernels
NBLK=8192 timeout 900 ncu --set full --import-source on --clock-control none -k regex:lz4_compress_wide -s 2 -c 1 -o $O/compress_full python tools/probe.py > $O/probe_under_ncu.log 2>&1
NBLK=8192 timeout 900 ncu --set full --import-source on --clock-control none -k regex:lz4_decompress_fast -s 2 -c 1 -o $O/decompress_fast_full python tools/probe.py > /dev/null 2>&1
timeout 600 python tools/corpus_sweep.py > $O/corpus_sweep.json 2> $O/corpus_sweep.err; tail -c 300 $O/corpus_sweep.json
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -W ignore::DeprecationWarning -k "not jni and not sweep and not past_4gb" > $O/sanitizer_memcheck.txt 2>&1; tail -4 $O/sanitizer_memcheck.txt | cut -c1-200
ls -la $O
