ulimit -c 0
timeout 900 python -m pytest tests -m gpu -x -q -W ignore::DeprecationWarning -k "decompress or frame or factory or self_roundtrip" > gpurun_out/p46.log 2>&1; tail -3 gpurun_out/p46.log | cut -c1-400
echo "== main (win 512)"; VARIANTS=13:0:3:0 timeout 600 python tools/probe.py 2>&1 | grep -E "decompress"
NBLK=512 BS=4194304 VARIANTS=12:0:3:0 timeout 600 python tools/probe.py 2>&1 | grep -E "decompress"
for v in w256 m10; do echo "== $v"; B200LZ4_SO=variants/libb200lz4_$v.so VARIANTS=13:0:3:0 timeout 600 python tools/probe.py 2>&1 | grep -E "decompress"; done
