#!/bin/bash
# First GPU-box visit of round 2 (round 1 ended with its GPU budget spent, so everything written after the last
# measurement is queued here):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/r02_first_call.sh'
# 1. the GPU parity suite (incl. the regression tests that have only run on the CPU emulator so far: fast-decoder
#    stream end, failed-pipeline drain, context reuse across threads, frames sharded by frame)
# 2. both bench arms on the current build -> the round's first BENCH line
# 3. the unmeasured second HC design (B200_EXPERIMENTAL: config 4 with b200lz4_hc_algo = 2) next to the default
# 4. end-to-end chunk-size sweep of the host pipeline (B200LZ4_CHUNK_MB)
# 5. launch list under ncu for the bench command
# 6. A/B (speed + instruction counts) of the two experimental variants of the fast compressor (B200_V3_RUNS, + B200_V3_SPLIT)
# Outputs under gpurun_out/r02a/ (copy what is to be judged into profiles/).
ulimit -c 0
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -W ignore::DeprecationWarning > $O/gpu_tests.log 2>&1; tail -3 $O/gpu_tests.log | cut -c1-300
B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests -m gpu -q -W ignore::DeprecationWarning -k "hc" > $O/gpu_tests_experimental.log 2>&1; tail -2 $O/gpu_tests_experimental.log | cut -c1-300
timeout 900 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference.err; tail -c 300 $O/bench_reference_arm.json
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err; tail -c 600 $O/bench_full.json
# 6. A/B of the run-start parser (B200_V3_RUNS=1: same bytes, fewer measurements; DESIGN.md round-2 plan), three corpora
bash tools/build_variants.sh runs:"-DB200_V3_RUNS=1" split:"-DB200_V3_RUNS=1 -DB200_V3_SPLIT=1" > $O/variant_build.log 2>&1
for mp in 0.5 0.8 0.2; do
  for so in lz4-java_b200/libb200lz4.so variants/libb200lz4_runs.so variants/libb200lz4_split.so; do
    echo "== MP=$mp $so"; MP=$mp B200LZ4_TEST_SO=$so NBLK=16384 VARIANTS=13:0:3:0,12:0:3:0 timeout 300 python tools/probe.py 2>&1 | tail -4 | cut -c1-300
  done
done > $O/runs_ab.log 2>&1; cat $O/runs_ab.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_issue_stalled_barrier.ratio
for so in lz4-java_b200/libb200lz4.so variants/libb200lz4_runs.so variants/libb200lz4_split.so; do
  B200LZ4_TEST_SO=$so NBLK=8192 VARIANTS=13:0:3:0 timeout 600 ncu --metrics $M --clock-control none -k regex:lz4_compress_fast3 -s 2 -c 1 --csv --log-file $O/runs_ab_$(basename $so .so).csv python tools/probe.py > /dev/null 2>&1
done
B200_EXPERIMENTAL=1 timeout 1500 python tools/bench_configs.py > $O/secondary_configs.log 2>&1; tail -12 $O/secondary_configs.log | cut -c1-300
for mb in 64 128 512; do B200LZ4_CHUNK_MB=$mb timeout 300 python tools/e2e_probe.py 2>&1 | head -3 | cut -c1-300; done > $O/e2e_chunk_sweep.log 2>&1; cat $O/e2e_chunk_sweep.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/bench_under_ncu.log 2>&1
ls -la $O

# Multi-GPU follow-up (separate call, N GPUs of one box):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'NBLK=65536 python tools/e2e_multi_probe.py > gpurun_out/e2e_multi.log 2>&1; \
#       python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_2gpu.json'
