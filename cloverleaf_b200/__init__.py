"""cloverleaf_b200 -- B200-native CloverLeaf hydro kernel layer.

The product is `libclover_b200.so` (hand-written sm_100a CUDA behind the
reference's `*_kernel_c_` C-ABI, see include/clover_b200.h) plus
`libclover_driver.so`, the host driver that restates the Fortran call sequence.
This Python package is only the loader / plumbing around those two libraries.
There is NO CPU fallback: `load_b200()` raises if the CUDA library is missing.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
# $CLOVER_B200_LIB: another build of the same library (A/B runs of compile-time variants, profiles/); default in-tree
LIB_B200 = os.environ.get("CLOVER_B200_LIB") or os.path.join(HERE, "libclover_b200.so")
LIB_DRIVER = os.path.join(HERE, "libclover_driver.so")
DECK_DIR = os.path.join(HERE, "decks")

_b200 = None


def load_b200():
    """dlopen the CUDA library (RTLD_GLOBAL so the driver's dlopen sees the same instance)."""
    global _b200
    if _b200 is None:
        if not os.path.exists(LIB_B200):
            raise RuntimeError(
                "libclover_b200.so is not built (run `python -m cloverleaf_b200.build`); "
                "there is no CPU fallback for the hot path")
        _b200 = ctypes.CDLL(LIB_B200, mode=ctypes.RTLD_GLOBAL)
    return _b200


from .driver import Driver, deck_text, KERNEL_SYMBOLS, EXTENSION_SYMBOLS  # noqa: E402,F401
