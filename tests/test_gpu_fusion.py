"""Deferred execution + cross-kernel fusion (csrc/fuse.cu) must be unobservable: every one of the 15 exchangeable
fields -- halos included -- and every dt are bit-identical with fusion on, with fusion off, and with the oracle,
at any point where the host looks (odd and even step counts: both advection sweep orders; mid-run downloads)."""
import ctypes

import numpy as np
import pytest

import cloverleaf_b200
from cloverleaf_b200.driver import Driver, deck_text
from conftest import ORACLE_PORT

pytestmark = pytest.mark.gpu

FIELDS = ("density0", "energy0", "xvel0", "yvel0", "pressure", "viscosity", "density1", "energy1", "xvel1",
          "yvel1", "vol_flux_x", "vol_flux_y", "mass_flux_x", "mass_flux_y", "soundspeed")


def _deck(nx, ny):
    return deck_text("clover_bm_short.in").replace("x_cells=960", "x_cells=%d" % nx).replace(
        "y_cells=960", "y_cells=%d" % ny)


def _set_fusion(lib, on):
    v = ctypes.c_int(1 if on else 0)
    lib.clover_b200_set_fusion_(ctypes.byref(v))


def _set_tma(lib, on):
    v = ctypes.c_int(1 if on else 0)
    lib.clover_b200_set_tma_(ctypes.byref(v))


def _fields(lib, d):
    out = {}
    for f in FIELDS:
        p = d._L.clover_driver_field(d._h, 0, f.encode())
        lib.clover_b200_download_(ctypes.c_void_p(p))
        out[f] = d.field(f).copy()
    return out


@pytest.fixture()
def fresh(b200):
    b200.clover_b200_invalidate_()
    _set_fusion(b200, True)
    _set_tma(b200, True)
    yield b200
    _set_fusion(b200, True)
    _set_tma(b200, True)
    b200.clover_b200_invalidate_()


@pytest.mark.parametrize("nx,ny,steps", [(64, 48, 7), (64, 48, 8), (250, 130, 21), (33, 2, 5), (2, 40, 6), (1, 1, 3), (3, 1, 3),
                                         (130, 97, 30)])
def test_fused_equals_unfused_equals_oracle(fresh, nx, ny, steps):
    deck = _deck(nx, ny)
    o = Driver(deck, ORACLE_PORT, end_step=steps); o.run()
    runs = {}
    # fused + TMA tile staging (the production path), fused with the register/L2-prefetch kernels, call by call
    for mode, (fuse, tma) in dict(tma=(True, True), fused=(True, False), unfused=(False, False)).items():
        fresh.clover_b200_invalidate_()
        _set_fusion(fresh, fuse)
        _set_tma(fresh, tma)
        d = Driver(deck, cloverleaf_b200.LIB_B200, end_step=steps); d.run()
        runs[mode] = (d.dts().copy(), _fields(fresh, d))
        d.close()
    for mode in runs:
        assert np.array_equal(o.dts(), runs[mode][0]), mode + " dt"
    for f in FIELDS:
        ref = o.field(f)
        for mode in runs:
            assert np.array_equal(ref, runs[mode][1][f]), mode + " " + f


def test_mid_run_downloads_do_not_disturb(fresh):
    """Looking at the device state between steps (sync hook of visit.f90) forces lazy copies and partial
    queues to materialise; the run must continue bit-identically."""
    deck = _deck(96, 80)
    o = Driver(deck, ORACLE_PORT, end_step=12); o.start()
    d = Driver(deck, cloverleaf_b200.LIB_B200, end_step=12); d.start()
    for _ in range(12):
        o.run(1); d.run(1)
        got = _fields(fresh, d)
        for f in FIELDS:
            assert np.array_equal(o.field(f), got[f]), f
    assert np.array_equal(o.dts(), d.dts())


def test_fusion_reduces_launches(fresh):
    deck = _deck(64, 64)
    n = ctypes.c_longlong(0)
    counts = {}
    for fuse in (True, False):
        fresh.clover_b200_invalidate_()
        _set_fusion(fresh, fuse)
        d = Driver(deck, cloverleaf_b200.LIB_B200, end_step=10); d.start()
        fresh.clover_b200_launch_count_(ctypes.byref(n)); a = n.value
        d.run(9)
        fresh.clover_b200_launch_count_(ctypes.byref(n)); counts[fuse] = n.value - a
        d.close()
    assert counts[True] < counts[False]
