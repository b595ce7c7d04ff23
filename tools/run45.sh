export NBLK=8192 VARIANTS=13:0:3:0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_decompress_fast -s 2 -c 1 -o gpurun_out/prof_decomp_r1k python tools/probe.py > gpurun_out/ncu_d.log 2>&1
tail -2 gpurun_out/ncu_d.log
