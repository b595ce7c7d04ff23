ulimit -c 0
set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --blocks 262144 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print({k:d[k] for k in ('value','n_gpus','compress_gibs','decompress_gibs','ratio','e2e','gpu_launches','clocks')})"
tail -5 gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --ref-blocks 8192 | cut -c1-300
