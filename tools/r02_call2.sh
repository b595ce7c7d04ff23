#!/bin/bash
# Round 2, GPU call 2: first run of the wide-chunk compressor (algo 5) next to algo 3, three corpora, + ncu counters.
ulimit -c 0
O=gpurun_out/r02c; mkdir -p $O
for mp in 0.5 0.8 0.2; do
  echo "== MP=$mp"; COMPRESS_ONLY=1 MP=$mp NBLK=16384 VARIANTS=13:22:5:0,13:12:5:0,13:42:5:0 timeout 300 python tools/probe.py 2>&1 | tail -7 | cut -c1-200
done > $O/wide_ab.log 2>&1; cat $O/wide_ab.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_wait.ratio,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
for v in 22; do
COMPRESS_ONLY=1 NBLK=8192 VARIANTS=13:$v:5:0 timeout 600 ncu --metrics $M --clock-control none -k regex:lz4_compress_wide -s 2 -c 1 --csv --log-file $O/wide_$v.csv python tools/probe.py > /dev/null 2>&1
grep -v "^==" $O/wide_$v.csv | cut -d, -f13- | cut -c1-150
done
COMPRESS_ONLY=1 NBLK=8192 VARIANTS=13:22:5:0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:lz4_compress_wide -s 2 -c 1 -o $O/wide_22_full python tools/probe.py > /dev/null 2>&1
ls -la $O
