set -x
timeout 900 python bench.py --blocks 131072 --steps 3 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -c 3000 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
for i in 1 2 3; do timeout 300 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/pytest_loop$i.log 2>&1; tail -1 gpurun_out/pytest_loop$i.log; done
