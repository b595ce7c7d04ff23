#!/bin/bash
# Round 2, GPU call 1 (trimmed from r02_first_call.sh to fit the 180-minute budget):
#   A/B of the fast-compress parser variants (speed + ncu instruction counts), first run of the second HC design,
#   re-run of the corpus sweep / P=0.80 diagnosis whose round-1 profile is stale.
ulimit -c 0
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
for mp in 0.5 0.8 0.2; do
  for so in lz4-java_b200/libb200lz4.so variants/libb200lz4_runs.so variants/libb200lz4_split.so; do
    echo "== MP=$mp $so"; COMPRESS_ONLY=1 MP=$mp B200LZ4_TEST_SO=$so NBLK=16384 VARIANTS=13:0:3:0 timeout 300 python tools/probe.py 2>&1 | tail -2 | cut -c1-300
  done
done > $O/runs_ab.log 2>&1; cat $O/runs_ab.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_wait.ratio
for so in lz4-java_b200/libb200lz4.so variants/libb200lz4_runs.so variants/libb200lz4_split.so; do
  COMPRESS_ONLY=1 B200LZ4_TEST_SO=$so NBLK=8192 VARIANTS=13:0:3:0 timeout 600 ncu --metrics $M --clock-control none -k regex:lz4_compress_fast3 -s 2 -c 1 --csv --log-file $O/runs_ab_$(basename $so .so).csv python tools/probe.py > /dev/null 2>&1
  tail -8 $O/runs_ab_$(basename $so .so).csv | cut -d, -f5,13- | cut -c1-200
done
B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests -m gpu -q -W ignore::DeprecationWarning -k "hc" > $O/gpu_tests_experimental.log 2>&1; tail -2 $O/gpu_tests_experimental.log | cut -c1-300
ONLY=config4 HC_NBLK=1024 B200_EXPERIMENTAL=1 timeout 900 python tools/bench_configs.py > $O/config4.log 2>&1; tail -25 $O/config4.log | cut -c1-300
timeout 600 python tools/corpus_sweep.py > $O/corpus_sweep.json 2> $O/corpus_sweep.err; tail -c 1500 $O/corpus_sweep.json
timeout 600 python tools/diag_p80.py > $O/diag_p80.log 2>&1; tail -5 $O/diag_p80.log | cut -c1-300
ls -la $O
