"""Generate the golden traces in this directory from the REFERENCE's own C kernels.

Runs oracle/_ref/libclover_ref_c.so (CloverLeaf_ref/kernels/*_kernel_c.c compiled where they lie,
-O3 -fopenmp -ffp-contract=off, OMP_NUM_THREADS=1 -- the configuration that reproduces the
reference's Intel-IEEE golden kinetic energies to the last printed digit) under the host driver and
records, for each case, every step's dt and every field_summary row at full precision.

Only runnable where /root/reference exists (this container); the JSON files are committed and are
what the GPU box checks against.   python tests/golden/make_golden.py [--big]
"""
import json
import os
import sys

# --threads N: dt (a min) and every field are independent of the OpenMP thread count (elementwise kernels);
# only the summed quantities move in the last digits (sum order), far inside the 1e-10 bar.  The long/huge
# cases below are generated with all cores; the small ones keep 1 thread.
_thr = "1"
if "--threads" in sys.argv:
    _thr = sys.argv[sys.argv.index("--threads") + 1]
os.environ["OMP_NUM_THREADS"] = _thr
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from cloverleaf_b200.driver import Driver, deck_text  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "libclover_ref_c.so")


def shrink(deck, n):
    return deck.replace("x_cells=960", "x_cells=%d" % n).replace("y_cells=960", "y_cells=%d" % n)


def exact_summary(d):
    """field_summary_kernel_c.c:52-88 evaluated on the reference run's own fields with the per-cell terms formed
    exactly as the kernel forms them (same operations, same order) but SUMMED in 80-bit pairwise arithmetic.  The
    reference accumulates serially in fp64, which at 7680^2 / 15360^2 cells carries ~2e-10 / ~1e-9 relative
    rounding of its own (its mass reads 28.000000005 for an exact 28); this is the sum a 1e-10 bar can be held to."""
    import ctypes
    import numpy as np
    from cloverleaf_b200.driver import FIELD_SHAPES
    info = d.chunk_info(0)
    nx, ny = info["x_max"], info["y_max"]

    def view(name):  # no copy: 15360^2 fields are 1.9 GB each
        ex, ey = FIELD_SHAPES[name]
        w, h = nx + 4 + ex, ny + 4 + ey
        p = d._L.clover_driver_field(d._h, 0, name.encode())
        return np.frombuffer((ctypes.c_double * (w * h)).from_address(p), dtype=np.float64).reshape(h, w)

    VOL, RHO, EN, PR, U, V = (view(n) for n in ("volume", "density0", "energy0", "pressure", "xvel0", "yvel0"))
    ld = np.longdouble
    acc = dict(volume=ld(0), mass=ld(0), ie=ld(0), ke=ld(0), pressure=ld(0))
    B = 256
    for k0 in range(2, 2 + ny, B):  # array row index of cell k is k+1; cells 1..ny -> rows 2..ny+1
        k1 = min(k0 + B, 2 + ny)
        sl = (slice(k0, k1), slice(2, 2 + nx))
        vol, rho, en, pr = VOL[sl], RHO[sl], EN[sl], PR[sl]
        vsq = np.zeros_like(vol)
        for kv in (0, 1):          # field_summary_kernel_c.c:74-79: kv outer, jv inner
            for jv in (0, 1):
                uu = U[k0 + kv:k1 + kv, 2 + jv:2 + nx + jv]
                vv = V[k0 + kv:k1 + kv, 2 + jv:2 + nx + jv]
                vsq = vsq + 0.25 * (uu * uu + vv * vv)
        mass = vol * rho
        acc["volume"] += np.sum(vol, dtype=ld)
        acc["mass"] += np.sum(mass, dtype=ld)
        acc["ie"] += np.sum(mass * en, dtype=ld)
        acc["ke"] += np.sum(mass * 0.5 * vsq, dtype=ld)
        acc["pressure"] += np.sum(vol * pr, dtype=ld)
    out = dict(step=float(d.step), **{k: float(v) for k, v in acc.items()})
    out["pressure"] = out["pressure"] / out["volume"]  # the driver reports press/vol (field_summary.f90:129)
    out["density"] = out["mass"] / out["volume"]
    out["total"] = out["ie"] + out["ke"]
    return out


def record(name, deck, nchunks=1, end_step=None, fields=False, exact=False):
    d = Driver(deck, REF, nchunks=nchunks, end_step=end_step)
    ex = []
    if exact:
        d.start()                  # start.f90:145 has just done the initial field_summary (ideal_gas included)
        ex.append(exact_summary(d))
    d.run()
    if exact:
        ex.append(exact_summary(d))  # the final field_summary ran ideal_gas on the final state
    G = dict(deck=deck, nchunks=nchunks, end_step=end_step, steps=d.step, dt=d.dts().tolist(),
             summaries=d.summaries(), summaries_exact=ex,
             source="oracle/_ref/libclover_ref_c.so (reference C kernels), %s thread(s)" % os.environ["OMP_NUM_THREADS"])
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(G, f, indent=0)
    print(name, "steps", d.step, "ke", repr(d.summaries()[-1]["ke"]))


def main_only(which):
    """The long (tp3, tp5) and huge (bm64, bm256) cases of BASELINE.json's configs, one at a time:
    python tests/golden/make_golden.py --threads 8 --only tp3_bm_960_full"""
    cases = {
        "tp3_bm_960_full": lambda: record("tp3_bm_960_full", deck_text("clover_bm.in")),
        "tp5_bm16_3840_full": lambda: record("tp5_bm16_3840_full", deck_text("clover_bm16.in")),
        "bm64_short_7680_first10": lambda: record("bm64_short_7680_first10", deck_text("clover_bm64_short.in"), end_step=10, exact=True),
        "bm256_short_15360_first10": lambda: record("bm256_short_15360_first10", deck_text("clover_bm256_short.in"), end_step=10, exact=True),
    }
    cases[which]()


if __name__ == "__main__":
    if "--only" in sys.argv:
        main_only(sys.argv[sys.argv.index("--only") + 1])
        sys.exit(0)
    bm = deck_text("clover_bm_short.in")
    record("tp1", deck_text("clover_tp1.in"))
    record("bm_short_96", shrink(bm, 96))
    record("multichunk_2x2_96", shrink(bm, 96), nchunks=4)
    record("bm_short_960_first10", bm, end_step=10)
    record("bm_short_960_full", bm)
    if "--big" in sys.argv:
        record("bm16_short_3840_full", deck_text("clover_bm16_short.in"))
