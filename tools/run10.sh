ulimit -c 0
export LD_PRELOAD=$(gcc -print-file-name=libasan.so)
export ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0:abort_on_error=0:halt_on_error=1:print_legend=0
export B200LZ4_SO=$PWD/lz4-java_b200/libb200lz4_asan.so
for i in 1 2 3; do
  timeout 600 python -X faulthandler -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning -k "decompress" > gpurun_out/p10_$i.log 2>&1
  echo "iter $i rc=$?"
  grep -m1 -A30 "ERROR: AddressSanitizer" gpurun_out/p10_$i.log | cut -c1-200 | head -45
  tail -2 gpurun_out/p10_$i.log | cut -c1-200
done
