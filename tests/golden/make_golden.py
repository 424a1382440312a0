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


def record(name, deck, nchunks=1, end_step=None, fields=False):
    d = Driver(deck, REF, nchunks=nchunks, end_step=end_step)
    d.run()
    G = dict(deck=deck, nchunks=nchunks, end_step=end_step, steps=d.step, dt=d.dts().tolist(),
             summaries=d.summaries(),
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
        "bm64_short_7680_first10": lambda: record("bm64_short_7680_first10", deck_text("clover_bm64_short.in"), end_step=10),
        "bm256_short_15360_first10": lambda: record("bm256_short_15360_first10", deck_text("clover_bm256_short.in"), end_step=10),
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
