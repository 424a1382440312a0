import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import cloverleaf_b200
from cloverleaf_b200.driver import Driver, deck_text
ORACLE_PORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "libclover_oracle.so")
FIELDS = ("density0", "energy0", "xvel0", "yvel0", "pressure", "viscosity", "density1", "energy1", "xvel1",
          "yvel1", "vol_flux_x", "vol_flux_y", "mass_flux_x", "mass_flux_y", "soundspeed")
lib = cloverleaf_b200.load_b200()
def deck(nx, ny):
    return deck_text("clover_bm_short.in").replace("x_cells=960", "x_cells=%d" % nx).replace("y_cells=960", "y_cells=%d" % ny)
def setf(on):
    v = ctypes.c_int(1 if on else 0); lib.clover_b200_set_fusion_(ctypes.byref(v))
def fields(d):
    out = {}
    for f in FIELDS:
        p = d._L.clover_driver_field(d._h, 0, f.encode())
        lib.clover_b200_download_(ctypes.c_void_p(p))
        out[f] = d.field(f).copy()
    return out
for nx, ny, steps in [(33, 2, 1), (33, 2, 2), (33, 2, 5), (2, 40, 6), (33, 3, 5), (33, 4, 5), (1, 1, 3), (3, 1, 3)]:
    o = Driver(deck(nx, ny), ORACLE_PORT, end_step=steps); o.run()
    for fuse in (True, False):
        lib.clover_b200_invalidate_(); setf(fuse)
        d = Driver(deck(nx, ny), cloverleaf_b200.LIB_B200, end_step=steps); d.run()
        got = fields(d)
        print(nx, ny, steps, "fuse" if fuse else "nofuse", "dt ok" if np.array_equal(o.dts(), d.dts()) else "DT DIFF")
        for f in FIELDS:
            ref = o.field(f)
            if not np.array_equal(ref, got[f]):
                bad = np.argwhere(ref != got[f])
                print("   ", f, ref.shape, "nbad", len(bad), "first", bad[:6].tolist(), "ref", [ref[tuple(b)] for b in bad[:3]], "got", [got[f][tuple(b)] for b in bad[:3]])
        d.close()
setf(True)
