"""bench.py's per-launch byte table must know every compute launch the library can time (a missing name made the
first TMA bench run die with StopIteration), and the fused launches' survey passes must add up to SURVEY.md's 107."""
import os
import re

from conftest import ROOT


def _launch_names():
    names = set()
    csrc = os.path.join(ROOT, "cloverleaf_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cuh")):
            src = open(os.path.join(csrc, f)).read()
            for m in re.finditer(r'LaunchScope\s+ls\((.*?)\);', src):
                names.update(re.findall(r'"([a-z0-9_]+)"', m.group(1)))
    return names


def test_every_compute_launch_has_a_byte_count():
    import bench
    not_streaming = {  # boundary / set-up / bookkeeping launches: no array-pass figure
        "update_halo", "update_halo_seq", "halo_pack", "halo_unpack", "halo_exchange_p2p", "allreduce_p2p",
        "reset_field_swap", "lazy_copy", "initialise_chunk_1d", "initialise_chunk_2d", "generate_chunk"}
    names = _launch_names()
    assert {"timestep_tma", "pdv_predict_tma", "lagrange_correct_tma", "advec_cell_x_tma", "advec_mom_y_tma"} <= names
    missing = names - not_streaming - set(bench.KERNEL_PASSES)
    assert not missing, "bench.KERNEL_PASSES lacks %s" % sorted(missing)


def test_fused_step_adds_up_to_the_survey_figure():
    import bench
    P = bench.KERNEL_PASSES
    step = ["timestep_tma", "pdv_predict_tma", "lagrange_correct_tma", "advec_cell_x_tma", "advec_cell_y_tma",
            "advec_mom_x_tma", "advec_mom_y_tma"]
    survey = sum(P[k][1] for k in step) + P["reset_field"][1]
    assert survey == 107 and survey * 8 == bench.ALG_BYTES_PER_CELL_STEP
    own = sum(P[k][0] for k in step)
    assert own == 64  # 512 B per cell-update actually moved (the sound speed is never stored)
