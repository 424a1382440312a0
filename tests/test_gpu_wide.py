"""Chunks wider than ~4096 cells are walked down 4096-column bands (tile_order, csrc/runtime.cu) instead of row by
row.  The order must not change a bit: a 4400 x 20 chunk (two bands for every tile width in use) against the oracle,
all 15 exchangeable fields incl. halos, dt at every step."""
import ctypes

import numpy as np
import pytest

import cloverleaf_b200
from cloverleaf_b200.driver import Driver, deck_text
from conftest import ORACLE_PORT

pytestmark = pytest.mark.gpu

FIELDS = ("density0", "energy0", "xvel0", "yvel0", "pressure", "viscosity", "density1", "energy1", "xvel1",
          "yvel1", "vol_flux_x", "vol_flux_y", "mass_flux_x", "mass_flux_y", "soundspeed")


def test_banded_tile_order_is_bit_identical(b200):
    nx, ny, steps = 4400, 20, 4
    deck = deck_text("clover_bm_short.in").replace("x_cells=960", "x_cells=%d" % nx).replace(
        "y_cells=960", "y_cells=%d" % ny)
    o = Driver(deck, ORACLE_PORT, end_step=steps)
    o.run()
    b200.clover_b200_invalidate_()
    d = Driver(deck, cloverleaf_b200.LIB_B200, end_step=steps)
    d.run()
    assert np.array_equal(o.dts(), d.dts())
    for f in FIELDS:
        p = d._L.clover_driver_field(d._h, 0, f.encode())
        b200.clover_b200_download_(ctypes.c_void_p(p))
        assert np.array_equal(o.field(f), d.field(f)), f
    d.close()
    b200.clover_b200_invalidate_()
