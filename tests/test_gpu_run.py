"""Whole-run parity on the GPU: the host driver (C++ restatement of the Fortran call sequence) runs
libclover_b200.so in RESIDENT mode and is compared with

  * the committed golden traces of the reference's own C kernels (tests/golden/*.json), and
  * the oracle run live on the same deck.

north_star bar: field_summary within 1e-10 relative at every summary step, dt step for step.
What we actually demand: dt BIT-IDENTICAL at every step and final fields bit-identical (the kernels
keep the reference's evaluation order, -fmad=false); summaries within 1e-12 relative (the device sum
is a fixed tree, the CPU sum is serial).
"""
import json
import os

import numpy as np
import pytest

import cloverleaf_b200
from cloverleaf_b200.driver import Driver, deck_text
from conftest import GOLDEN, ORACLE_PORT

pytestmark = pytest.mark.gpu

# The north_star's bar.  The golden sums come from a SERIAL CPU accumulation that itself carries
# ~1e-11 relative rounding at 960^2 (the committed volume reads 99.99999999880 for an exact 100);
# the device tree sum is the more accurate of the two (100.00000000000003).
SUM_TOL = 1e-10


def _check_against(G_dt, G_sum, d, tol=SUM_TOL):
    dts = d.dts().tolist()
    assert len(dts) == len(G_dt)
    first_bad = next((i for i, (a, b) in enumerate(zip(dts, G_dt)) if a != b), None)
    assert first_bad is None, "dt differs first at step %d: %r vs %r" % (
        first_bad + 1, dts[first_bad], G_dt[first_bad])
    got = d.summaries()
    assert len(got) == len(G_sum)
    for a, b in zip(got, G_sum):
        assert a["step"] == b["step"] and a["time"] == b["time"]
        for k in ("volume", "mass", "density", "pressure", "ie", "ke", "total"):
            assert abs(a[k] - b[k]) <= tol * max(abs(b[k]), 1e-300) or (a[k] == b[k]), (b["step"], k, a[k], b[k])


@pytest.fixture()
def fresh(b200):
    b200.clover_b200_invalidate_()
    yield b200
    b200.clover_b200_invalidate_()


@pytest.mark.parametrize("name", ["tp1", "bm_short_96", "bm_short_960_first10", "bm_short_960_full"])
def test_run_matches_reference_golden(fresh, name):
    G = json.load(open(os.path.join(GOLDEN, name + ".json")))
    d = Driver(G["deck"], cloverleaf_b200.LIB_B200, end_step=G.get("end_step"))
    d.run()
    _check_against(G["dt"], G["summaries"], d)


def test_golden_ke_constants(fresh):
    """The reference's own pins (field_summary.f90:139-140): tp1 and tp2 kinetic energies."""
    d = Driver("clover_tp1.in", cloverleaf_b200.LIB_B200)
    d.run()
    assert abs(d.summaries()[-1]["ke"] / 1.82280367310258 - 1.0) < 1e-13
    fresh.clover_b200_invalidate_()
    d = Driver("clover_bm_short.in", cloverleaf_b200.LIB_B200)
    d.run()
    assert d.step == 87
    assert abs(d.summaries()[-1]["ke"] / 1.19316898756307 - 1.0) < 1e-12


def test_final_fields_bit_identical_to_oracle(fresh):
    deck = deck_text("clover_bm_short.in").replace("x_cells=960", "x_cells=250").replace("y_cells=960", "y_cells=130")
    a = Driver(deck, ORACLE_PORT); a.run()
    b = Driver(deck, cloverleaf_b200.LIB_B200); b.run()
    assert np.array_equal(a.dts(), b.dts())
    import ctypes
    for f in ("density0", "energy0", "xvel0", "yvel0", "pressure", "viscosity", "density1", "energy1", "xvel1",
              "yvel1", "vol_flux_x", "vol_flux_y", "mass_flux_x", "mass_flux_y", "soundspeed"):
        # sync_to_host needs register_chunk (comm_mode 1); in comm_mode 0 download by address
        p = b._L.clover_driver_field(b._h, 0, f.encode())
        fresh.clover_b200_download_(ctypes.c_void_p(p))
        assert np.array_equal(a.field(f), b.field(f)), f


@pytest.mark.parametrize("nchunks", [2, 4, 8])
def test_decomposition_independence_on_one_gpu(fresh, nchunks):
    """N chunks in one process on one GPU (exchange through the ABI pack/unpack kernels and host
    buffers, as MPI would): dt bit-identical to the 1-chunk run (README.md:103-109 claim)."""
    deck = deck_text("clover_bm_short.in").replace("x_cells=960", "x_cells=120").replace("y_cells=960", "y_cells=96")
    one = Driver(deck, cloverleaf_b200.LIB_B200, end_step=40); one.run()
    dt1, s1 = one.dts(), one.summaries()
    fresh.clover_b200_invalidate_()
    many = Driver(deck, cloverleaf_b200.LIB_B200, nchunks=nchunks, end_step=40); many.run()
    assert np.array_equal(dt1, many.dts())
    for a, b in zip(s1, many.summaries()):
        for k in ("volume", "mass", "ie", "ke", "pressure"):
            assert abs(a[k] - b[k]) <= 1e-12 * max(abs(a[k]), 1e-300)


def _run_against_golden(name, deck, end_step=None):
    d = Driver(deck, cloverleaf_b200.LIB_B200, end_step=end_step)
    d.run()
    path = os.path.join(GOLDEN, name + ".json")
    assert os.path.exists(path), "golden trace %s missing (tests/golden/make_golden.py --only %s)" % (path, name)
    G = json.load(open(path))
    _check_against(G["dt"], G["summaries"], d)
    return d


def test_bm16_short_full_run_tp4(fresh):
    """BASELINE's headline configuration, clover_bm16_short (3840^2, all 87 steps): dt bit-identical at every
    step and every summary row within 1e-10 of the reference's C kernels (committed trace), the reference's own
    pin tp4 (field_summary.f90:142), and the size-independent invariants (README.md:246-251)."""
    d = _run_against_golden("bm16_short_3840_full", "clover_bm16_short.in")
    assert d.step == 87
    s = d.summaries()
    assert abs(s[-1]["ke"] / 0.307475452287895 - 1.0) < 1e-10
    assert abs(s[0]["volume"] - 100.0) < 1e-9 and abs(s[-1]["volume"] - 100.0) < 1e-9
    assert abs(s[-1]["mass"] / s[0]["mass"] - 1.0) < 1e-12
    assert abs(s[-1]["total"] / s[0]["total"] - 1.0) < 1e-3


def test_bm_full_run_tp3(fresh):
    """clover_bm.in (960^2, 2955 steps): the reference's pin tp3 (field_summary.f90:141) and the committed
    step-for-step trace of the reference's C kernels (all 2955 dt values, 296 summary rows)."""
    d = _run_against_golden("tp3_bm_960_full", "clover_bm.in")
    assert d.step == 2955
    assert abs(d.summaries()[-1]["ke"] / 2.58984003503994 - 1.0) < 1e-10


def test_bm16_full_run_tp5(fresh):
    """clover_bm16.in (3840^2, 2955 steps): the reference's pin tp5 (field_summary.f90:143); the committed trace of
    the reference's C kernels when it is there (40 CPU-minutes to generate)."""
    d = Driver("clover_bm16.in", cloverleaf_b200.LIB_B200)
    d.run()
    assert d.step == 2955
    assert abs(d.summaries()[-1]["ke"] / 4.85350315783719 - 1.0) < 1e-10
    path = os.path.join(GOLDEN, "tp5_bm16_3840_full.json")
    if os.path.exists(path):
        G = json.load(open(path))
        _check_against(G["dt"], G["summaries"], d)


@pytest.mark.parametrize("name,deck", [("bm64_short_7680_first10", "clover_bm64_short.in"),
                                       ("bm256_short_15360_first10", "clover_bm256_short.in")])
def test_bm64_bm256_first_steps(fresh, name, deck):
    """The two large BASELINE configurations (7680^2, 15360^2) on one GPU: the first 10 steps (dt bit-identical,
    two summary rows) against the committed traces of the reference's C kernels.  15360^2 exercises the banded
    tile order (chunks wider than 4096 columns)."""
    d = Driver(deck, cloverleaf_b200.LIB_B200, end_step=10)
    d.run()
    G = json.load(open(os.path.join(GOLDEN, name + ".json")))
    # The reference accumulates its sums serially in fp64; at these sizes that alone carries 2e-10 .. 1e-9 relative
    # rounding (its mass reads 28.000000005 for an exact 28), so against its printed sums only 1e-8 can be asked --
    _check_against(G["dt"], G["summaries"], d, tol=1e-8)
    # -- and the 1e-10 bar (1e-12 here) is held against the same per-cell terms of the reference run's own fields
    # summed in 80-bit pairwise arithmetic (tests/golden/make_golden.py: exact_summary).
    s = d.summaries()  # rows: start-up (step 0), the periodic one of step 10, the final one (step 10 again)
    assert [int(r["step"]) for r in G["summaries_exact"]] == [0, 10] and [int(r["step"]) for r in s] == [0, 10, 10]
    for a, b in zip([s[0], s[1], s[2]], [G["summaries_exact"][0], G["summaries_exact"][1], G["summaries_exact"][1]]):
        for k in ("volume", "mass", "density", "pressure", "ie", "ke", "total"):
            assert abs(a[k] - b[k]) <= 1e-12 * max(abs(b[k]), 1e-300), (b["step"], k, a[k], b[k])
    assert d.step == 10
    assert abs(s[-1]["volume"] - 100.0) < 1e-9
    assert abs(s[-1]["mass"] / s[0]["mass"] - 1.0) < 1e-12
