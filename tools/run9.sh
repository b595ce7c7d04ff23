ulimit -c 0
for t in test_decompress_safe_exact test_decompress_fast_exact test_decompress_safe_malformed_codes test_decompress_fast_malformed test_xxhash_batches test_compact_host; do
  for i in 1 2 3; do
    MALLOC_PERTURB_=165 timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning -k "$t" > gpurun_out/p9_${t}_$i.log 2>&1
    echo "$t iter $i rc=$?"
  done
done
# same but without PERTURB and without the reference lib (port checker)
