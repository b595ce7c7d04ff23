#!/bin/bash
# Round 2, second (short) visit to an 8-GPU box: what the host side can move at all (8 GPUs copying both ways at once, node-local
# and not), and the one-process multi-GPU leg with node-local host buffers.
ulimit -c 0
O=gpurun_out/r02p; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 300 $TR tools/dma_ceiling.py > $O/dma_ceiling_numa.json 2> $O/dma1.err; tail -c 600 $O/dma_ceiling_numa.json
timeout 300 $TR tools/dma_ceiling.py --no-numa > $O/dma_ceiling_nonuma.json 2> $O/dma2.err; tail -c 600 $O/dma_ceiling_nonuma.json
timeout 600 $TR bench.py --gpus 8 --no-cpu --steps 3 --blocks 131072 --only none > $O/bench_8gpu_e2e_multi.json 2> $O/bench3.err; tail -c 900 $O/bench_8gpu_e2e_multi.json; tail -3 $O/bench3.err
