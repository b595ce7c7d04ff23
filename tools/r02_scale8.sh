#!/bin/bash
# Round 2, 8-GPU box: the bench line at N=8 (ranks pinned to their GPU's NUMA node), and the same end-to-end leg without
# the pinning; topology for the record.
ulimit -c 0
O=gpurun_out/r02n; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; (numactl -H || lscpu | grep -i numa) >> $O/topo.txt 2>&1; nproc >> $O/topo.txt; cat /sys/fs/cgroup/cpu.max >> $O/topo.txt 2>/dev/null
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 1500 $TR bench.py --gpus 8 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; tail -c 300 $O/bench_8gpu.json; tail -2 $O/bench_8gpu.err
timeout 600 $TR bench.py --gpus 8 --no-numa --no-secondary --no-cpu --steps 3 --blocks 131072 > $O/bench_8gpu_nonuma.json 2> $O/bench_8gpu_nonuma.err; tail -c 300 $O/bench_8gpu_nonuma.json
timeout 600 $TR bench.py --gpus 8 --no-secondary --no-cpu --steps 3 --blocks 131072 > $O/bench_8gpu_numa_short.json 2> $O/bench_8gpu_numa_short.err; tail -c 300 $O/bench_8gpu_numa_short.json
timeout 300 python bench.py --impl reference --gpus 8 > $O/bench_ref_8.json 2>/dev/null; tail -c 300 $O/bench_ref_8.json
