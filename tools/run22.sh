export NBLK=8192 VARIANTS=13:0:3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_compress_fast3 -s 2 -c 1 -o gpurun_out/prof_compress_r1f python tools/probe.py > gpurun_out/ncu_c.log 2>&1
