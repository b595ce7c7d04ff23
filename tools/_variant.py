"""Development helper for the tools in this directory: B200LZ4_TEST_SO=<path> points the package's loader at another
build of libb200lz4.so (tools/build_variants.sh) before first use.  Import it before touching lz4java_b200."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_alt = os.environ.get("B200LZ4_TEST_SO")
if _alt:
    import lz4java_b200._native as _N
    _N.SO_PATH = os.path.abspath(_alt)
