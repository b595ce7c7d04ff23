export NBLK=512 BS=4194304 VARIANTS=12:0:3:0
for v in old d8; do echo "== $v"; B200LZ4_SO=variants/libb200lz4_$v.so timeout 600 python tools/probe.py 2>&1 | grep -E "^compress|decompress"; done
echo "== main"; timeout 600 python tools/probe.py 2>&1 | grep -E "decompress"
