ulimit -c 0
set -x
# 2-GPU weak scaling through torchrun, small batch
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --blocks 262144 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 1800 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --ref-blocks 8192 | cut -c1-300
