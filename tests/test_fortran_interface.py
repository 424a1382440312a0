"""include/clover_b200_kernels.f90 (the ISO_C_BINDING spelling of the ABI, north_star) is generated from the same
argument table the parity tests call the library through; no Fortran compiler exists here, so it is checked as text:
it is up to date, it binds every kernel symbol, and every name it binds is exported by libclover_b200.so."""
import ctypes
import os
import re
import subprocess
import sys

from conftest import ROOT

F90 = os.path.join(ROOT, "include", "clover_b200_kernels.f90")


def test_module_is_up_to_date():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "include", "gen_fortran_interface.py")],
                         capture_output=True, text=True, check=True).stdout
    assert out == open(F90).read(), "re-run: python include/gen_fortran_interface.py > include/clover_b200_kernels.f90"


def test_module_binds_the_whole_kernel_abi_and_only_exported_symbols():
    import cloverleaf_b200
    from cloverleaf_b200 import abi
    src = open(F90).read()
    bound = re.findall(r"BIND\(C, NAME='([a-z0-9_]+)'\)", src)
    assert len(bound) == len(set(bound))
    assert set(abi.KERNEL_SYMBOLS) <= set(bound)
    lib = ctypes.CDLL(cloverleaf_b200.LIB_B200)  # loading needs no GPU
    for sym in bound:
        getattr(lib, sym)
    # argument counts of the kernel entry points match the table
    for sym, spec in abi.KERNELS.items():
        m = re.search(r"SUBROUTINE %s\((.*?)\)\s*&\s*\n\s*BIND\(C, NAME='%s'\)" % (sym[:-1], sym), src, re.S)
        assert m, sym
        args = [a.strip() for a in m.group(1).replace("&", " ").replace("\n", " ").split(",")]
        assert args == [a for a, _ in spec], sym
