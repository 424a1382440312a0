"""Seeded, physically plausible inputs for every kernel entry point, and an A/B runner.

A "case" is (kernel symbol, scalar kwargs).  `make_state` builds one random chunk state (all 25
fields + geometry); `run_case` calls one kernel of one backend on a private copy of that state and
returns the arrays the kernel may have written.
"""
import numpy as np

from cloverleaf_b200 import abi

WORK = ["work%d" % i for i in range(1, 8)]
FIELD_IDS = ["density0", "density1", "energy0", "energy1", "pressure", "viscosity", "soundspeed",
             "xvel0", "xvel1", "yvel0", "yvel1", "vol_flux_x", "vol_flux_y", "mass_flux_x", "mass_flux_y"]
FIELD_TYPE = dict(density0="C", density1="C", energy0="C", energy1="C", pressure="C", viscosity="C",
                  soundspeed="C", xvel0="V", xvel1="V", yvel0="V", yvel1="V", vol_flux_x="X",
                  vol_flux_y="Y", mass_flux_x="X", mass_flux_y="Y", volume="C", xarea="X", yarea="Y")


def make_state(nx, ny, seed=0):
    rng = np.random.default_rng(seed)
    dx, dy = 10.0 / nx, 7.0 / ny
    S = {"nx": nx, "ny": ny}

    def r(code, lo, hi):
        return np.ascontiguousarray(rng.uniform(lo, hi, abi.shape(code, nx, ny)))

    S["density0"] = r("C", 0.5, 1.5)
    S["density1"] = r("C", 0.5, 1.5)
    S["energy0"] = r("C", 1.0, 3.0)
    S["energy1"] = r("C", 1.0, 3.0)
    S["pressure"] = r("C", 0.2, 1.8)
    S["viscosity"] = r("C", 0.0, 0.2)
    S["viscosity"][S["viscosity"] < 0.05] = 0.0
    S["soundspeed"] = r("C", 0.5, 1.5)
    for v in ("xvel0", "xvel1", "yvel0", "yvel1"):
        S[v] = r("V", -0.5, 0.5)
    S["volume"] = np.full(abi.shape("C", nx, ny), dx * dy)
    S["xarea"] = np.full(abi.shape("X", nx, ny), dy)
    S["yarea"] = np.full(abi.shape("Y", nx, ny), dx)
    S["vol_flux_x"] = r("X", -0.1, 0.1) * dx * dy
    S["vol_flux_y"] = r("Y", -0.1, 0.1) * dx * dy
    S["mass_flux_x"] = r("X", -0.1, 0.1) * dx * dy
    S["mass_flux_y"] = r("Y", -0.1, 0.1) * dx * dy
    # exact zeros exercise the `> 0.0` / `< 0.0` branch edges of the upwind selection
    for f in ("vol_flux_x", "vol_flux_y", "mass_flux_x", "mass_flux_y"):
        S[f][rng.uniform(size=S[f].shape) < 0.02] = 0.0
    for w in WORK:
        S[w] = np.zeros(abi.shape("W", nx, ny))
    S["cellx"] = (np.arange(-1, nx + 3) - 0.5) * dx
    S["celly"] = (np.arange(-1, ny + 3) - 0.5) * dy
    S["vertexx"] = (np.arange(-1, nx + 4) - 1.0) * dx
    S["vertexy"] = (np.arange(-1, ny + 4) - 1.0) * dy
    # slightly non-uniform widths so that width ratios are not all exactly 1
    S["celldx"] = dx * (1.0 + 0.01 * rng.uniform(-1, 1, nx + 4))
    S["celldy"] = dy * (1.0 + 0.01 * rng.uniform(-1, 1, ny + 4))
    S["vertexdx"] = dx * (1.0 + 0.01 * rng.uniform(-1, 1, nx + 5))
    S["vertexdy"] = dy * (1.0 + 0.01 * rng.uniform(-1, 1, ny + 5))
    return S


def copy_state(S):
    return {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in S.items()}


def _base_kwargs(S):
    return dict(x_min=1, x_max=S["nx"], y_min=1, y_max=S["ny"])


def run_case(lib, S0, kernel, **scalars):
    """Run `kernel` from `lib` on a copy of state S0; return (state copy, dict of scalar outputs)."""
    S = copy_state(S0)
    kw = _base_kwargs(S)
    out = {}
    spec = abi.KERNELS[kernel]
    alias = scalars.pop("alias", {})
    for arg, code in spec:
        if arg in kw:
            continue
        if arg in scalars:
            v = scalars[arg]
            if code in ("i", "d") and isinstance(v, str) and v == "out":
                v = np.zeros(1, dtype=np.int32 if code == "i" else np.float64)
                out[arg] = v
            kw[arg] = v
        elif alias.get(arg, arg) in S:
            kw[arg] = S[alias.get(arg, arg)]
        else:
            raise KeyError("%s: no value for %s" % (kernel, arg))
    abi.call(lib, kernel, **kw)
    return S, {k: v[0] for k, v in out.items()}


DT_PARAMS = dict(g_small=1.0e-16, g_big=1.0e21, dtmin=1.0e-7, dtc_safe=0.7, dtu_safe=0.5, dtv_safe=0.5,
                 dtdiv_safe=0.7)


def kernel_cases():
    """(id, kernel, scalars) for every hot-path entry point and every mode it is called in."""
    cases = [
        ("ideal_gas", "ideal_gas_kernel_c_", dict(alias=dict(density="density0", energy="energy0"))),
        ("ideal_gas_predict", "ideal_gas_kernel_c_", dict(alias=dict(density="density1", energy="energy1"))),
        ("viscosity", "viscosity_kernel_c_", {}),
        ("calc_dt", "calc_dt_kernel_c_", dict(DT_PARAMS, dt_min_val="out", dtl_control="out", xl_pos="out",
                                               yl_pos="out", jldt="out", kldt="out", small="out",
                                               alias=dict(dt_min="work1"))),
        ("pdv_predict", "pdv_kernel_c_", dict(prdct=0, dt=0.013, alias=dict(volume_change="work1"))),
        ("pdv_correct", "pdv_kernel_c_", dict(prdct=1, dt=0.013, alias=dict(volume_change="work1"))),
        ("revert", "revert_kernel_c_", {}),
        ("accelerate", "accelerate_kernel_c_", dict(dt=0.013)),
        ("flux_calc", "flux_calc_kernel_c_", dict(dt=0.013)),
        ("reset_field", "reset_field_kernel_c_", {}),
        ("field_summary", "field_summary_kernel_c_", dict(vol="out", mass="out", ie="out", ke="out", press="out")),
    ]
    for d in (1, 2):
        for s in (1, 2):
            cases.append(("advec_cell_dir%d_sweep%d" % (d, s), "advec_cell_kernel_c_", dict(dir=d, sweep_number=s)))
    for d in (1, 2):
        for s in (1, 2):
            for w, vel in ((1, "xvel1"), (2, "yvel1")):
                cases.append(("advec_mom_dir%d_sweep%d_vel%d" % (d, s, w), "advec_mom_kernel_c_",
                              dict(which_vel=w, sweep_number=s, direction=d, alias=dict(vel1=vel))))
    return cases


def halo_cases():
    """update_halo: every combination of external faces x depth, all 15 fields requested."""
    out = []
    for depth in (1, 2):
        for mask in range(1, 16):
            nb = np.array([(-1 if mask & (1 << f) else 7) for f in range(4)], dtype=np.int32)
            out.append(("halo_d%d_ext%x" % (depth, mask), depth, nb))
    return out
