ulimit -c 0
timeout 900 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/p44.log 2>&1; tail -4 gpurun_out/p44.log | cut -c1-400
for n in 2048 4096 8192 16384; do for below in 0 1000000; do echo "n=$n batch_below=$below"; NBLK=$n DEC_BELOW=$below VARIANTS=13:0:3:0 timeout 600 python tools/probe.py 2>&1 | grep -E "decompress"; done; done
