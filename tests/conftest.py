import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_PORT = os.path.join(ROOT, "oracle", "libclover_oracle.so")
ORACLE_REF = os.path.join(ROOT, "oracle", "_ref", "libclover_ref_c.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the numpy reductions in the oracle comparisons and the OpenMP oracle must be deterministic
    os.environ.setdefault("OMP_NUM_THREADS", "1")


def _ensure_built():
    from cloverleaf_b200 import build
    build.build_driver()
    if not os.path.exists(ORACLE_PORT) or (
            os.path.isdir("/root/reference") and not os.path.exists(ORACLE_REF)):
        build.build_oracle()
    import cloverleaf_b200
    if not os.path.exists(cloverleaf_b200.LIB_B200):
        build.build_b200()


@pytest.fixture(scope="session", autouse=True)
def built():
    _ensure_built()


@pytest.fixture(scope="session")
def oracle_lib(built):
    """The plain-C restatement (oracle/clover_oracle.c) -- the checker, never the product."""
    return ctypes.CDLL(ORACLE_PORT)


@pytest.fixture(scope="session")
def ref_lib(built):
    """The reference's own C kernels compiled from /root/reference (oracle/_ref), if present."""
    if not os.path.exists(ORACLE_REF):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return ctypes.CDLL(ORACLE_REF)


def have_gpu():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20)
        return out.returncode == 0 and "GPU" in out.stdout
    except Exception:
        return False


@pytest.fixture(scope="session")
def b200(built):
    """The CUDA library.  GPU tests FAIL (not skip) if it cannot be loaded on a GPU box."""
    import cloverleaf_b200
    lib = cloverleaf_b200.load_b200()
    dev = ctypes.c_int(0)
    lib.clover_b200_init_(ctypes.byref(dev))
    return lib
