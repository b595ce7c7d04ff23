set -x
timeout 600 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning > gpurun_out/pytest7.log 2>&1; tail -3 gpurun_out/pytest7.log
VARIANTS=13:0,12:0 timeout 600 python tools/probe.py 2>&1 | tail -12
export NBLK=8192 VARIANTS=13:0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_compress_fast -s 2 -c 1 -o gpurun_out/prof_compress_r1c python tools/probe.py > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_decompress_safe -s 2 -c 1 -o gpurun_out/prof_decsafe_r1c python tools/probe.py > gpurun_out/ncu_d.log 2>&1
