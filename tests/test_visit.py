"""visit.f90:25-180 in the host driver: index file + one ASCII VTK rectilinear-grid file per chunk and dump, values in
Fortran E12.4.  Checked with the oracle backend (CPU); the GPU backend goes through the same code plus the D2H of the
six dumped fields (tests/test_gpu_zvisit.py)."""
import os

import numpy as np

from cloverleaf_b200.driver import Driver, deck_text
from conftest import ORACLE_PORT


def _deck(nx, ny):
    return deck_text("clover_bm_short.in").replace("x_cells=960", "x_cells=%d" % nx).replace(
        "y_cells=960", "y_cells=%d" % ny)


def _fortran_e12_4(v):
    """What gfortran/ifort print for '(e12.4)'."""
    if v == 0.0:
        return "  0.0000E+00"
    e = int(np.floor(np.log10(abs(v)))) + 1
    m = round(abs(v) / 10.0 ** e * 1e4)
    if m >= 10000:
        m, e = 1000, e + 1
    s = "%s0.%04dE%s%02d" % ("-" if v < 0 else "", m, "-" if e < 0 else "+", abs(e))
    return s.rjust(12)


def test_e12_4_known_values():
    for v, want in [(1.0, "  0.1000E+01"), (0.2, "  0.2000E+00"), (2.5, "  0.2500E+01"), (-0.05, " -0.5000E-01"),
                    (12345.678, "  0.1235E+05"), (0.99996, "  0.1000E+01"), (1e-9, "  0.1000E-08")]:
        assert _fortran_e12_4(v) == want


def _parse_vtk(path):
    lines = open(path).read().split("\n")
    assert lines[:4] == ["# vtk DataFile Version 3.0", "vtk output", "ASCII", "DATASET RECTILINEAR_GRID"]
    out, i = {}, 4
    nxv, nyv = int(lines[i][10:22]), int(lines[i][22:34]); i += 1
    assert lines[i] == "X_COORDINATES %5d double" % nxv; i += 1
    out["x"] = lines[i:i + nxv]; i += nxv
    assert lines[i] == "Y_COORDINATES %5d double" % nyv; i += 1
    out["y"] = lines[i:i + nyv]; i += nyv
    assert lines[i:i + 2] == ["Z_COORDINATES 1 double", "0"]; i += 2
    nc = (nxv - 1) * (nyv - 1)
    assert lines[i] == "CELL_DATA %20d" % nc and lines[i + 1] == "FIELD FieldData 4"; i += 2
    for name in ("density", "energy", "pressure", "viscosity"):
        assert lines[i] == "%s 1 %20d double" % (name, nc); i += 1
        out[name] = lines[i:i + nc]; i += nc
    assert lines[i] == "POINT_DATA %20d" % (nxv * nyv) and lines[i + 1] == "FIELD FieldData 2"; i += 2
    for name in ("x_vel", "y_vel"):
        assert lines[i] == "%s 1 %20d double" % (name, nxv * nyv); i += 1
        out[name] = lines[i:i + nxv * nyv]; i += nxv * nyv
    assert lines[i:] == [""]
    return nxv, nyv, out


def test_visit_files_single_chunk(tmp_path):
    nx, ny, steps = 24, 16, 12
    d = Driver(_deck(nx, ny), ORACLE_PORT, end_step=steps)
    d.set_visit(tmp_path, 5)
    d.run()
    index = open(tmp_path / "clover.visit").read().split("\n")
    assert index[0] == "!NBLOCKS     1"
    # start.f90:145 (step 0), hydro.f90:73-75 (steps 5, 10), hydro.f90:88 (final step 12)
    assert index[1:-1] == ["clover.00000.00001.%05d.vtk" % s for s in (0, 5, 10, 12)]
    nxv, nyv, v = _parse_vtk(tmp_path / "clover.00000.00001.00012.vtk")
    assert (nxv, nyv) == (nx + 1, ny + 1)
    rho = d.field("density0")[2:2 + ny, 2:2 + nx].ravel()
    assert v["density"] == [_fortran_e12_4(x) for x in rho]
    p = d.field("pressure")[2:2 + ny, 2:2 + nx].ravel()
    assert v["pressure"] == [_fortran_e12_4(x) for x in p]
    q = d.field("viscosity")[2:2 + ny, 2:2 + nx].ravel()
    assert v["viscosity"] == [_fortran_e12_4(x if x > 1e-8 else 0.0) for x in q]
    u = d.field("xvel0")[2:3 + ny, 2:3 + nx].ravel()
    assert v["x_vel"] == [_fortran_e12_4(x if abs(x) > 1e-8 else 0.0) for x in u]
    assert v["x"][0] == _fortran_e12_4(0.0) and v["x"][-1] == _fortran_e12_4(10.0)
    assert any(s.strip().startswith("-") for s in v["y_vel"]) or any(s != "  0.0000E+00" for s in v["y_vel"])


def test_visit_files_four_chunks(tmp_path):
    d = Driver(_deck(20, 20), ORACLE_PORT, nchunks=4, end_step=3)
    d.set_visit(tmp_path, 0)  # frequency from the deck (0): no dumps at all
    d.run()
    assert not os.path.exists(tmp_path / "clover.visit")
    d = Driver(_deck(20, 20), ORACLE_PORT, nchunks=4, end_step=3)
    d.set_visit(tmp_path, 2)
    d.run()
    index = open(tmp_path / "clover.visit").read().split("\n")
    assert index[0] == "!NBLOCKS     4"
    assert len(index) - 2 == 4 * 3  # dumps at steps 0, 2, 3
    for task in range(4):
        nxv, nyv, _ = _parse_vtk(tmp_path / ("clover.%05d.00001.00003.vtk" % task))
        assert (nxv, nyv) == (11, 11)
