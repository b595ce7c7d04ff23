set -x
timeout 900 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning 2>&1 | tail -25
timeout 600 python tools/probe.py 2>&1 | tail -14
export NBLK=4096 VARIANTS=13:0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_compress_fast -s 2 -c 1 -o gpurun_out/prof_compress_r1b python tools/probe.py > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_decompress_safe -s 2 -c 1 -o gpurun_out/prof_decsafe_r1b python tools/probe.py > gpurun_out/ncu_d.log 2>&1
