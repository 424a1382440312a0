"""CPU tests that PIN the oracle (oracle/clover_oracle.c) -- no GPU needed.

 1. bit-for-bit against the reference's own C kernels compiled from /root/reference (oracle/_ref),
    per entry point on seeded random inputs (work arrays included) and on whole runs;
 2. against the reference's golden kinetic energies (CloverLeaf_ref/field_summary.f90:139-143);
 3. against the committed traces in tests/golden/ (written by tests/golden/make_golden.py from
    oracle/_ref), which is what travels to the GPU box where /root/reference does not exist.
"""
import json
import os

import numpy as np
import pytest

import kernel_cases as kc
from cloverleaf_b200 import abi
from cloverleaf_b200.driver import Driver
from conftest import GOLDEN, ORACLE_PORT, ORACLE_REF

SIZES = [(13, 9), (32, 17), (5, 40)]


def _assert_same_state(A, B, skip=()):
    for k, v in A.items():
        if isinstance(v, np.ndarray) and k not in skip:
            assert np.array_equal(v, B[k], equal_nan=True), "array %s differs" % k


@pytest.mark.parametrize("nx,ny", SIZES)
@pytest.mark.parametrize("case", kc.kernel_cases(), ids=lambda c: c[0])
def test_port_matches_reference_kernels(oracle_lib, ref_lib, case, nx, ny):
    name, kernel, scalars = case
    S0 = kc.make_state(nx, ny, seed=nx * 100 + ny)
    A, oa = kc.run_case(ref_lib, S0, kernel, **dict(scalars))
    B, ob = kc.run_case(oracle_lib, S0, kernel, **dict(scalars))
    _assert_same_state(A, B)
    for k in oa:
        if k in ("xl_pos", "yl_pos", "small"):
            continue  # passed through / never written by the reference (calc_dt_kernel_c.c:161-162)
        assert oa[k] == ob[k], (name, k, oa[k], ob[k])


@pytest.mark.parametrize("case", kc.halo_cases(), ids=lambda c: c[0])
def test_port_matches_reference_update_halo(oracle_lib, ref_lib, case):
    _, depth, nb = case
    S0 = kc.make_state(11, 7, seed=3)
    fields = np.ones(15, dtype=np.int32)
    kw = dict(chunk_neighbours=nb, tile_neighbours=np.full(4, -1, dtype=np.int32), fields=fields, depth=depth)
    A, _ = kc.run_case(ref_lib, S0, "update_halo_kernel_c_", **kw)
    B, _ = kc.run_case(oracle_lib, S0, "update_halo_kernel_c_", **kw)
    _assert_same_state(A, B)


@pytest.mark.parametrize("face", ["left", "right", "bottom", "top"])
@pytest.mark.parametrize("depth", [1, 2])
def test_port_matches_reference_pack_unpack(oracle_lib, ref_lib, face, depth):
    nx, ny = 9, 6
    S0 = kc.make_state(nx, ny, seed=5)
    for fname, ftype in (("density0", abi.CELL_DATA), ("xvel0", abi.VERTEX_DATA),
                         ("vol_flux_x", abi.X_FACE_DATA), ("mass_flux_y", abi.Y_FACE_DATA)):
        res = []
        for lib in (ref_lib, oracle_lib):
            S = kc.copy_state(S0)
            buf = np.zeros(10 * 2 * (max(nx, ny) + 5))
            common = dict(x_min=1, x_max=nx, y_min=1, y_max=ny, field=S[fname], buffer=buf,
                          cell_data=abi.CELL_DATA, vertex_data=abi.VERTEX_DATA, x_face_data=abi.X_FACE_DATA,
                          y_face_data=abi.Y_FACE_DATA, depth=depth, field_type=ftype, buffer_offset=7)
            abi.call(lib, "clover_pack_message_%s_c_" % face, **common)
            packed = buf.copy()
            buf[:] = np.arange(buf.size) + 0.5
            abi.call(lib, "clover_unpack_message_%s_c_" % face, **common)
            res.append((packed, S[fname].copy()))
        assert np.array_equal(res[0][0], res[1][0])
        assert np.array_equal(res[0][1], res[1][1])


def test_port_matches_reference_setup_kernels(oracle_lib, ref_lib):
    nx, ny = 12, 10
    states = dict(
        number_of_states=4,
        state_density=np.array([0.2, 1.0, 0.7, 0.5]), state_energy=np.array([1.0, 2.5, 1.5, 3.0]),
        state_xvel=np.array([0.0, 0.1, -0.2, 0.3]), state_yvel=np.array([0.0, -0.1, 0.2, 0.05]),
        state_xmin=np.array([0.0, 0.0, 6.0, 4.0]), state_xmax=np.array([0.0, 5.0, 0.0, 0.0]),
        state_ymin=np.array([0.0, 0.0, 4.0, 4.0]), state_ymax=np.array([0.0, 2.0, 0.0, 0.0]),
        state_radius=np.array([0.0, 0.0, 2.0, 0.0]), state_geometry=np.array([1, 1, 2, 3], dtype=np.int32),
        g_rect=1, g_circ=2, g_point=3)
    res = []
    for lib in (ref_lib, oracle_lib):
        S = kc.make_state(nx, ny, seed=1)
        for k in ("vertexx", "vertexdx", "vertexy", "vertexdy", "cellx", "celldx", "celly", "celldy"):
            S[k] = np.zeros_like(S[k])
        for k in ("volume", "xarea", "yarea", "density0", "energy0", "xvel0", "yvel0"):
            S[k] = np.zeros_like(S[k])
        S, _ = kc.run_case(lib, S, "initialise_chunk_kernel_c_", min_x=0.0, min_y=0.0, dx=10.0 / nx, dy=10.0 / ny)
        S, _ = kc.run_case(lib, S, "generate_chunk_kernel_c_", **states)
        res.append(S)
    _assert_same_state(res[0], res[1])
    assert (res[1]["density0"] == 0.7).any() and (res[1]["density0"] == 1.0).any()


def _trace(d):
    return dict(dt=d.dts().tolist(), summaries=d.summaries())


def test_port_matches_reference_whole_run(ref_lib, oracle_lib):
    """87 steps of a 96x96 version of the benchmark deck: dt and summaries bit-identical."""
    deck = open(os.path.join(os.path.dirname(abi.__file__), "decks", "clover_bm_short.in")).read()
    deck = deck.replace("x_cells=960", "x_cells=96").replace("y_cells=960", "y_cells=96")
    a = Driver(deck, ORACLE_REF); a.run()
    b = Driver(deck, ORACLE_PORT); b.run()
    assert a.step == b.step == 87
    assert np.array_equal(a.dts(), b.dts())
    assert a.summaries() == b.summaries()
    for f in ("density0", "energy0", "xvel0", "yvel0"):
        assert np.array_equal(a.field(f), b.field(f))


def test_golden_kinetic_energy_tp1(oracle_lib):
    """Built-in 10x2 deck (initialise.f90:77-91): KE 1.82280367310258 (field_summary.f90:139)."""
    d = Driver("clover_tp1.in", ORACLE_PORT)
    d.run()
    assert d.step == 75
    ke = d.summaries()[-1]["ke"]
    assert abs(ke / 1.82280367310258 - 1.0) < 1e-14


@pytest.mark.parametrize("name", ["tp1", "bm_short_96", "bm_short_960_first10", "multichunk_2x2_96"])
def test_port_matches_golden_traces(oracle_lib, name):
    path = os.path.join(GOLDEN, name + ".json")
    G = json.load(open(path))
    d = Driver(G["deck"], ORACLE_PORT, nchunks=G.get("nchunks", 1), end_step=G.get("end_step"))
    d.run()
    assert d.dts().tolist() == G["dt"], "dt trace differs from the reference's"
    got = d.summaries()
    assert len(got) == len(G["summaries"])
    for a, b in zip(got, G["summaries"]):
        for k in b:
            if G.get("nchunks", 1) == 1:
                assert a[k] == b[k], (name, k)
            else:
                assert abs(a[k] - b[k]) <= 1e-13 * max(1.0, abs(b[k])), (name, k)


def test_golden_tp2_constant_is_in_fixture():
    """clover_bm_short (960^2, 87 steps): the committed reference run reproduces the reference's golden
    KE 1.19316898756307 (field_summary.f90:140) -- ties the fixtures to the reference's own pin."""
    G = json.load(open(os.path.join(GOLDEN, "bm_short_960_full.json")))
    assert G["summaries"][-1]["step"] == 87
    assert abs(G["summaries"][-1]["ke"] / 1.19316898756307 - 1.0) < 1e-13
