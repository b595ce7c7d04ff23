import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    # Development switch of the TEST harness (the package has none): run the suite against another build of the
    # library — compute-sanitizer / variant builds, or tests/simt's emulator build on a box without a GPU.
    alt = os.environ.get("B200LZ4_TEST_SO")
    if alt:
        import lz4java_b200._native as N
        N.SO_PATH = os.path.abspath(alt)


@pytest.fixture(scope="session")
def port():
    from oracle import oracle as O
    return O.Port()


@pytest.fixture(scope="session")
def ref():
    """The reference's own C (oracle/_ref).  Present wherever oracle/Makefile could build it or the
    prebuilt .so travelled; tests that need it are skipped otherwise."""
    from oracle import oracle as O
    try:
        return O.Ref()
    except (FileNotFoundError, OSError):
        pytest.skip("oracle/_ref/liblz4ref.so not available")


@pytest.fixture(scope="session")
def checker():
    """Best available CPU checker: reference build if present, else the pinned port."""
    from oracle import oracle as O
    return O.best_available()


@pytest.fixture(scope="session")
def b200():
    import lz4java_b200 as L
    return L
