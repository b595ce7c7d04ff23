set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v Deprecation | tail -15
export NBLK=4096
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_compress_fast -s 2 -c 1 -o gpurun_out/prof_compress_r1a python tools/probe.py > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_decompress_safe -s 2 -c 1 -o gpurun_out/prof_decsafe_r1a python tools/probe.py > gpurun_out/ncu_d.log 2>&1
tail -3 gpurun_out/ncu_c.log gpurun_out/ncu_d.log
ls -la gpurun_out
