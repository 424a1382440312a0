"""CPU tests of the boundary and of the host logic (no GPU, no compute calls into the CUDA library).

 * libclover_b200.so loads and exports every symbol include/clover_b200.h declares;
 * the Python ABI table (cloverleaf_b200/abi.py) agrees with the header;
 * the host driver: deck parsing, clover_decompose, the multi-chunk exchange emulation
   (decomposition independence of the reference C kernels, README.md:103-109) and the stop rules.
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import cloverleaf_b200
from cloverleaf_b200 import abi
from cloverleaf_b200.driver import Driver, deck_text
from conftest import ORACLE_PORT, ROOT

HEADER = os.path.join(ROOT, "include", "clover_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"\bvoid\s+([a-z0-9_]+_)\s*\(", text)


def test_header_declares_the_reference_entry_points():
    syms = header_symbols()
    for s in abi.KERNEL_SYMBOLS:
        assert s in syms, s
    assert len(abi.KERNEL_SYMBOLS) == 22  # 14 kernels + 8 pack/unpack, timer_c_ separately
    assert "timer_c_" in syms


def test_library_exports_every_declared_symbol():
    out = subprocess.run(["nm", "-D", "--defined-only", cloverleaf_b200.LIB_B200], capture_output=True, text=True)
    exported = {line.split()[-1] for line in out.stdout.splitlines() if line.strip()}
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, missing
    # and it dlopens without a GPU (no compute call is made here)
    lib = ctypes.CDLL(cloverleaf_b200.LIB_B200)
    for s in abi.KERNEL_SYMBOLS + abi.EXTENSION_SYMBOLS:
        assert hasattr(lib, s), s


def test_python_abi_table_matches_header_arity():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, spec in abi.KERNELS.items():
        m = re.search(r"\bvoid\s+%s\s*\(([^;]*)\)\s*;" % re.escape(name), text, flags=re.S)
        assert m, name
        nargs = len([a for a in m.group(1).split(",") if a.strip()])
        assert nargs == len(spec), (name, nargs, len(spec))


def test_no_cpu_fallback_in_product_package():
    """The product package must not import or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "cloverleaf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                if f == "build.py":
                    continue  # builds the checker (allowed), never loads it
                assert "clover_oracle" not in src and "libclover_ref" not in src, os.path.join(dirpath, f)


# ---- host driver ------------------------------------------------------------------------------------
def test_deck_parser_and_defaults():
    d = Driver("clover_bm16_short.in", ORACLE_PORT, end_step=0)
    g = d.grid()
    assert (g["x_cells"], g["y_cells"]) == (3840, 3840)
    with pytest.raises(RuntimeError):
        Driver("*clover\n x_cells=4\n*endclover\n", ORACLE_PORT)  # no states
    # `=` and `,` are blanks, `!` starts a comment, keywords are case-insensitive (parse.f90:160-184)
    deck = "*clover\n STATE 1 density 0.2, energy=1.0 ! background\n x_cells 12 ; y_cells 7\n y_cells=7\n end_step=2\n*endclover\n"
    d = Driver(deck, ORACLE_PORT)
    assert d.grid()["x_cells"] == 12 and d.grid()["y_cells"] == 7


@pytest.mark.parametrize("n,expect", [(1, (1, 1)), (2, (2, 1)), (4, (2, 2)), (8, (2, 4)), (6, (2, 3)), (3, (3, 1))])
def test_decompose_square_mesh(n, expect):
    """clover_decompose (clover.f90:127-196); SURVEY.md 8e: 2 -> 2x1, 4 -> 2x2, 8 -> 2(x) x 4(y)."""
    deck = deck_text("clover_bm_short.in").replace("x_cells=960", "x_cells=96").replace("y_cells=960", "y_cells=96")
    d = Driver(deck, ORACLE_PORT, nchunks=n, end_step=1)
    d.start()
    g = d.grid()
    assert (g["chunk_x"], g["chunk_y"]) == expect
    assert d.num_local_chunks() == n
    cells = 0
    for i in range(n):
        c = d.chunk_info(i)
        cells += c["x_max"] * c["y_max"]
        assert c["id"] == i + 1
        cx, cy = i % g["chunk_x"], i // g["chunk_x"]
        assert c["nb_left"] == (-1 if cx == 0 else i)
        assert c["nb_right"] == (-1 if cx == g["chunk_x"] - 1 else i + 2)
        assert c["nb_bottom"] == (-1 if cy == 0 else i + 1 - g["chunk_x"])
        assert c["nb_top"] == (-1 if cy == g["chunk_y"] - 1 else i + 1 + g["chunk_x"])
    assert cells == 96 * 96


def test_decompose_remainders_go_to_low_chunks():
    deck = deck_text("clover_bm_short.in").replace("x_cells=960", "x_cells=101").replace("y_cells=960", "y_cells=50")
    d = Driver(deck, ORACLE_PORT, nchunks=4, end_step=1)
    d.start()
    widths = sorted({d.chunk_info(i)["x_max"] for i in range(4)})
    assert sum(d.chunk_info(i)["x_max"] * d.chunk_info(i)["y_max"] for i in range(4)) == 101 * 50
    assert widths[-1] - widths[0] <= 1


@pytest.mark.parametrize("n", [2, 4, 8])
def test_multichunk_matches_single_chunk_bitwise(n):
    """Same deck, 1 chunk vs n chunks in one process (exchange through the C pack/unpack kernels):
    dt bit-identical at every step, fields bit-identical, sums equal to rounding."""
    deck = deck_text("clover_bm_short.in").replace("x_cells=960", "x_cells=60").replace("y_cells=960", "y_cells=44")
    a = Driver(deck, ORACLE_PORT, end_step=30); a.run()
    b = Driver(deck, ORACLE_PORT, nchunks=n, end_step=30); b.run()
    assert np.array_equal(a.dts(), b.dts())
    for f in ("density0", "energy0", "xvel0", "yvel0"):
        assert np.array_equal(a.global_field(f), b.global_field(f)), f
    for x, y in zip(a.summaries(), b.summaries()):
        for k in ("volume", "mass", "ie", "ke", "pressure"):
            assert abs(x[k] - y[k]) <= 1e-12 * max(abs(x[k]), 1e-300)


def test_stop_rules_and_summary_cadence():
    """hydro.f90:70-84: summary every summary_frequency steps, plus the final one; stop on end_step."""
    deck = deck_text("clover_bm_short.in").replace("x_cells=960", "x_cells=24").replace("y_cells=960", "y_cells=24")
    d = Driver(deck, ORACLE_PORT, end_step=25)
    assert d.run() == 25 and d.complete
    steps = [int(s["step"]) for s in d.summaries()]
    assert steps == [0, 10, 20, 25]
    d = Driver(deck, ORACLE_PORT, end_step=20)
    d.run()
    assert [int(s["step"]) for s in d.summaries()] == [0, 10, 20, 20]  # final summary repeats step 20
    # conservation (README.md:246-251)
    s = d.summaries()
    assert abs(s[-1]["mass"] / s[0]["mass"] - 1.0) < 1e-13 and abs(s[-1]["volume"] / s[0]["volume"] - 1.0) < 1e-13


def test_first_step_dt_rule():
    """timestep.f90:97 with dtold seeded by dtinit (start.f90:50): dt1 = min(dtlp, dtinit*dtrise, dtmax)."""
    deck = deck_text("clover_bm_short.in").replace("x_cells=960", "x_cells=24").replace("y_cells=960", "y_cells=24")
    d = Driver(deck, ORACLE_PORT, end_step=3)
    d.run()
    dts = d.dts()
    assert dts[0] <= 0.04 and dts[1] <= dts[0] * 1.5 + 1e-18 and dts[2] <= dts[1] * 1.5 + 1e-18
