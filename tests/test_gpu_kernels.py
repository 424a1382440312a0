"""Parity tests proper: every CUDA kernel, called through the C-ABI of libclover_b200.so with HOST
arrays (copy-in / copy-out mode), against the oracle on the same seeded inputs.

Bar: BIT-EXACT for every field a kernel writes (the library is built -fmad=false and keeps the
reference's evaluation order).  Only the sum reductions of field_summary are compared with a
tolerance (1e-11 relative, 10x inside the north_star's 1e-10: the device sum is a fixed tree, the CPU
sum is serial and carries ~n*eps of its own rounding); the calc_dt minimum is exact.
Work arrays are scratch in the reference and are not written by the device kernels, so they are
excluded from the comparison.
"""
import ctypes

import numpy as np
import pytest

import kernel_cases as kc
from cloverleaf_b200 import abi

pytestmark = pytest.mark.gpu

# ragged sizes: not multiples of the 32x8 tiles / 31-wide warp runs / 32-row column segments,
# a single-row and a single-column chunk, and one larger than any single block
SIZES = [(13, 9), (31, 33), (64, 32), (1, 40), (50, 1), (257, 131)]


@pytest.fixture(scope="module")
def cuda(b200):
    off = ctypes.c_int(0)
    b200.clover_b200_set_resident_(ctypes.byref(off))
    yield b200
    on = ctypes.c_int(1)
    b200.clover_b200_set_resident_(ctypes.byref(on))
    b200.clover_b200_invalidate_()


def _assert_fields_equal(A, B, what):
    for k, v in A.items():
        if isinstance(v, np.ndarray) and k not in kc.WORK:
            if not np.array_equal(v, B[k], equal_nan=True):
                bad = np.argwhere(v != B[k])
                raise AssertionError("%s: %s differs at %d points, first (k+1,j+1)=%s oracle=%r cuda=%r" % (
                    what, k, len(bad), bad[0], v[tuple(bad[0])], B[k][tuple(bad[0])]))


@pytest.mark.parametrize("nx,ny", SIZES)
@pytest.mark.parametrize("case", kc.kernel_cases(), ids=lambda c: c[0])
def test_kernel_matches_oracle(cuda, oracle_lib, case, nx, ny):
    name, kernel, scalars = case
    S0 = kc.make_state(nx, ny, seed=7 * nx + ny)
    A, oa = kc.run_case(oracle_lib, S0, kernel, **dict(scalars))
    B, ob = kc.run_case(cuda, S0, kernel, **dict(scalars))
    _assert_fields_equal(A, B, name)
    if name == "calc_dt":
        assert oa["dt_min_val"] == ob["dt_min_val"]
        assert ob["dtl_control"] == 1 and ob["jldt"] == 1 and ob["kldt"] == 1
    if name == "field_summary":
        for k in ("vol", "mass", "ie", "ke", "press"):
            assert abs(oa[k] - ob[k]) <= 1e-11 * abs(oa[k]), (k, oa[k], ob[k])


@pytest.mark.parametrize("case", kc.halo_cases(), ids=lambda c: c[0])
# chunks narrower than depth+1 take the four-launch sequential path (mirror sources inside the halo)
@pytest.mark.parametrize("nx,ny", [(11, 7), (40, 3), (2, 5), (9, 1), (1, 1)])
def test_update_halo_matches_oracle(cuda, oracle_lib, case, nx, ny):
    _, depth, nb = case
    S0 = kc.make_state(nx, ny, seed=3)
    kw = dict(chunk_neighbours=nb, tile_neighbours=np.full(4, -1, dtype=np.int32),
              fields=np.ones(15, dtype=np.int32), depth=depth)
    A, _ = kc.run_case(oracle_lib, S0, "update_halo_kernel_c_", **kw)
    B, _ = kc.run_case(cuda, S0, "update_halo_kernel_c_", **kw)
    _assert_fields_equal(A, B, "update_halo")


def test_update_halo_field_mask(cuda, oracle_lib):
    """Only requested fields are touched (the six per-step request lists of SURVEY.md 2.4)."""
    S0 = kc.make_state(17, 12, seed=11)
    nb = np.array([-1, 3, -1, -1], dtype=np.int32)
    for ids, depth in (([4, 2, 0, 7, 9], 1), ([5], 1), ([3, 1, 11, 12], 2), ([1, 3, 8, 10, 13, 14], 2)):
        fields = np.zeros(15, dtype=np.int32)
        fields[ids] = 1
        kw = dict(chunk_neighbours=nb, tile_neighbours=np.full(4, -1, dtype=np.int32), fields=fields, depth=depth)
        A, _ = kc.run_case(oracle_lib, S0, "update_halo_kernel_c_", **kw)
        B, _ = kc.run_case(cuda, S0, "update_halo_kernel_c_", **kw)
        _assert_fields_equal(A, B, "update_halo mask %s" % ids)


@pytest.mark.parametrize("face", ["left", "right", "bottom", "top"])
@pytest.mark.parametrize("depth", [1, 2])
def test_pack_unpack_matches_oracle(cuda, oracle_lib, face, depth):
    nx, ny = 37, 21
    S0 = kc.make_state(nx, ny, seed=5)
    for fname, ftype in (("density0", abi.CELL_DATA), ("xvel0", abi.VERTEX_DATA),
                         ("vol_flux_x", abi.X_FACE_DATA), ("mass_flux_y", abi.Y_FACE_DATA)):
        res = []
        for lib in (oracle_lib, cuda):
            S = kc.copy_state(S0)
            buf = np.zeros(10 * 2 * (max(nx, ny) + 5))
            common = dict(x_min=1, x_max=nx, y_min=1, y_max=ny, field=S[fname], buffer=buf,
                          cell_data=abi.CELL_DATA, vertex_data=abi.VERTEX_DATA, x_face_data=abi.X_FACE_DATA,
                          y_face_data=abi.Y_FACE_DATA, depth=depth, field_type=ftype, buffer_offset=2 * depth * 26)
            abi.call(lib, "clover_pack_message_%s_c_" % face, **common)
            packed = buf.copy()
            buf[:] = np.arange(buf.size) + 0.5
            abi.call(lib, "clover_unpack_message_%s_c_" % face, **common)
            res.append((packed, S[fname].copy()))
        assert np.array_equal(res[0][0], res[1][0]), (fname, "pack")
        assert np.array_equal(res[0][1], res[1][1]), (fname, "unpack")


def test_setup_kernels_match_oracle(cuda, oracle_lib):
    nx, ny = 45, 38
    states = dict(
        number_of_states=4,
        state_density=np.array([0.2, 1.0, 0.7, 0.5]), state_energy=np.array([1.0, 2.5, 1.5, 3.0]),
        state_xvel=np.array([0.0, 0.1, -0.2, 0.3]), state_yvel=np.array([0.0, -0.1, 0.2, 0.05]),
        state_xmin=np.array([0.0, 0.0, 6.0, 4.0]), state_xmax=np.array([0.0, 5.0, 0.0, 0.0]),
        state_ymin=np.array([0.0, 0.0, 4.0, 4.0]), state_ymax=np.array([0.0, 2.0, 0.0, 0.0]),
        state_radius=np.array([0.0, 0.0, 2.0, 0.0]), state_geometry=np.array([1, 1, 2, 3], dtype=np.int32),
        g_rect=1, g_circ=2, g_point=3)
    res = []
    for lib in (oracle_lib, cuda):
        S = kc.make_state(nx, ny, seed=1)
        for k in ("vertexx", "vertexdx", "vertexy", "vertexdy", "cellx", "celldx", "celly", "celldy",
                  "volume", "xarea", "yarea", "density0", "energy0", "xvel0", "yvel0"):
            S[k] = np.zeros_like(S[k])
        S, _ = kc.run_case(lib, S, "initialise_chunk_kernel_c_", min_x=0.0, min_y=0.0, dx=10.0 / nx, dy=10.0 / ny)
        S, _ = kc.run_case(lib, S, "generate_chunk_kernel_c_", **states)
        res.append(S)
    _assert_fields_equal(res[0], res[1], "initialise/generate")


def test_resident_mode_round_trip(b200, oracle_lib):
    """Resident mode: arrays are uploaded on first sight, stay on the device across calls, and come
    back with clover_b200_download_; chained kernels give the oracle's chained result."""
    on = ctypes.c_int(1)
    b200.clover_b200_set_resident_(ctypes.byref(on))
    b200.clover_b200_invalidate_()
    nx, ny = 70, 45
    S0 = kc.make_state(nx, ny, seed=21)
    A = kc.copy_state(S0)
    B = kc.copy_state(S0)

    def chain(lib, S):
        base = dict(x_min=1, x_max=nx, y_min=1, y_max=ny)
        abi.call(lib, "ideal_gas_kernel_c_", density=S["density0"], energy=S["energy0"], pressure=S["pressure"],
                 soundspeed=S["soundspeed"], **base)
        abi.call(lib, "viscosity_kernel_c_", celldx=S["celldx"], celldy=S["celldy"], density0=S["density0"],
                 pressure=S["pressure"], viscosity=S["viscosity"], xvel0=S["xvel0"], yvel0=S["yvel0"], **base)
        abi.call(lib, "accelerate_kernel_c_", dt=0.01, xarea=S["xarea"], yarea=S["yarea"], volume=S["volume"],
                 density0=S["density0"], pressure=S["pressure"], viscosity=S["viscosity"], xvel0=S["xvel0"],
                 yvel0=S["yvel0"], xvel1=S["xvel1"], yvel1=S["yvel1"], **base)

    chain(oracle_lib, A)
    before = B["xvel1"].copy()
    chain(b200, B)
    assert np.array_equal(B["xvel1"], before), "resident mode must not touch host arrays"
    for f in ("pressure", "soundspeed", "viscosity", "xvel1", "yvel1"):
        b200.clover_b200_download_(ctypes.c_void_p(B[f].ctypes.data))
        assert np.array_equal(A[f], B[f]), f
    b200.clover_b200_invalidate_()


def test_fast_math_matches_ieee(b200):
    """The branch-free div / rcp / sqrt sequences (common.cuh Math<false>) are bit-identical to the
    compiler's IEEE operators on 3 x 2^28 operands (random over 2^+-600, hydro-like magnitudes,
    adversarial mantissas, signed zeros); operands outside the guarded range are flagged, not wrong."""
    n, seed = ctypes.c_longlong(1 << 28), ctypes.c_longlong(20261017)
    mism, flagged, checked = ctypes.c_longlong(-1), ctypes.c_longlong(-1), ctypes.c_longlong(-1)
    b200.clover_b200_selftest_math_(ctypes.byref(n), ctypes.byref(seed), ctypes.byref(mism),
                                    ctypes.byref(flagged), ctypes.byref(checked))
    assert mism.value == 0, "%d mismatches in %d checked" % (mism.value, checked.value)
    assert checked.value > 0.9 * 3 * n.value, (checked.value, flagged.value)
