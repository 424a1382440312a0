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

os.environ["OMP_NUM_THREADS"] = "1"
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
             summaries=d.summaries(), source="oracle/_ref/libclover_ref_c.so (reference C kernels), 1 thread")
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(G, f, indent=0)
    print(name, "steps", d.step, "ke", repr(d.summaries()[-1]["ke"]))


if __name__ == "__main__":
    bm = deck_text("clover_bm_short.in")
    record("tp1", deck_text("clover_tp1.in"))
    record("bm_short_96", shrink(bm, 96))
    record("multichunk_2x2_96", shrink(bm, 96), nchunks=4)
    record("bm_short_960_first10", bm, end_step=10)
    record("bm_short_960_full", bm)
    if "--big" in sys.argv:
        record("bm16_short_3840_full", deck_text("clover_bm16_short.in"))
