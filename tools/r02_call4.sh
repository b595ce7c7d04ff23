#!/bin/bash
# Round 2: wide compressor, two- vs three-warp builds; ncu counters + full capture of the three-warp build.
ulimit -c 0
O=gpurun_out/r02g; mkdir -p $O
for mp in 0.5 0.8 0.2; do
  echo "== MP=$mp"; COMPRESS_ONLY=1 MP=$mp NBLK=16384 VARIANTS=${VARIANTS:-13:222:5:0,13:322:5:0,13:312:5:0} timeout 300 python tools/probe.py 2>&1 | tail -4 | cut -c1-200
done > $O/wide_ab.log 2>&1; cat $O/wide_ab.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
for v in ${NCU_V:-322}; do
COMPRESS_ONLY=1 NBLK=8192 VARIANTS=13:$v:5:0 timeout 600 ncu --metrics $M --clock-control none -k regex:lz4_compress_wide -s 2 -c 1 --csv --log-file $O/wide_$v.csv python tools/probe.py > /dev/null 2>&1
COMPRESS_ONLY=1 NBLK=8192 VARIANTS=13:$v:5:0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:lz4_compress_wide -s 2 -c 1 -o $O/wide_${v}_full python tools/probe.py > /dev/null 2>&1
done
ls -la $O
