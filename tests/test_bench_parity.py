"""bench.py's parity gate (the multi-GPU parity proof on the driver's record) and the trace summariser, on CPU."""
import json
import os
import subprocess
import sys

from conftest import GOLDEN, ROOT


def test_parity_check_accepts_the_golden_trace_and_rejects_a_flipped_bit():
    import bench
    G = json.load(open(os.path.join(GOLDEN, "bm16_short_3840_full.json")))
    ok = bench.parity_check("clover_bm16_short.in", G["dt"][:48], [r for r in G["summaries"] if r["step"] <= 48])
    assert ok["ok"] and ok["dt_bit_identical"] and ok["dt_steps_checked"] == 48 and ok["summary_rows_checked"] >= 4
    dt = list(G["dt"][:48])
    import numpy as np
    dt[30] = float(np.nextafter(dt[30], 1.0))  # one unit in the last place
    bad = bench.parity_check("clover_bm16_short.in", dt, [])
    assert bad["ok"] is False and bad["dt_bit_identical"] is False and bad["first_dt_mismatch_step"] == 31
    rows = [dict(r) for r in G["summaries"] if r["step"] <= 20]
    rows[-1]["ke"] *= 1.0 + 5e-10
    assert bench.parity_check("clover_bm16_short.in", G["dt"][:20], rows)["ok"] is False


def test_parity_check_uses_the_exact_sums_for_the_huge_decks():
    import bench
    G = json.load(open(os.path.join(GOLDEN, "bm256_short_15360_first10.json")))
    # the reference's own serial sums are ~1e-9 off at 15360^2: against them the 1e-10 gate would fail ...
    naive = [r for r in G["summaries"] if r["step"] in (0, 10)]
    exact = G["summaries_exact"]
    worst = max(abs(a[k] - b[k]) / abs(b[k]) for a, b in zip(naive[:2], exact) for k in ("mass", "ie", "volume"))
    assert worst > 1e-10
    # ... so the gate compares with the exact re-summation of the reference run's own fields
    assert bench.parity_check("clover_bm256_short.in", G["dt"], [dict(r, time=0.0) for r in exact])["ok"]


def test_every_baseline_deck_has_a_golden_trace():
    import bench
    for deck, name in bench.GOLDEN_FOR.items():
        assert os.path.exists(os.path.join(ROOT, "cloverleaf_b200", "decks", deck)), deck
        assert os.path.exists(os.path.join(GOLDEN, name)), name


def test_trace_summary_reads_a_committed_timeline():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "trace_summary.py"),
                          os.path.join(ROOT, "profiles", "r02_trace_8gpu_rank0.csv")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "halo_exchange_p2p" in out.stdout and "timestep_tma" in out.stdout
