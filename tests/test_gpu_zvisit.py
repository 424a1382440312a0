"""The visit dump through the CUDA backend (device -> host of the six dumped fields inside visit()) must produce the
same files, byte for byte, as the oracle backend."""
import pytest

import cloverleaf_b200
from cloverleaf_b200.driver import Driver, deck_text
from conftest import ORACLE_PORT

pytestmark = pytest.mark.gpu


def test_visit_dump_matches_oracle(b200, tmp_path):
    deck = deck_text("clover_bm_short.in").replace("x_cells=960", "x_cells=70").replace("y_cells=960", "y_cells=45")
    outs = {}
    for name, so in (("oracle", ORACLE_PORT), ("b200", cloverleaf_b200.LIB_B200)):
        b200.clover_b200_invalidate_()
        out = tmp_path / name
        out.mkdir()
        d = Driver(deck, so, end_step=11)
        d.set_visit(out, 4)
        d.run()
        d.close()
        outs[name] = {p.name: p.read_bytes() for p in sorted(out.iterdir())}
    assert sorted(outs["oracle"]) == sorted(outs["b200"]) and len(outs["oracle"]) == 1 + 4  # steps 0, 4, 8, 11
    for fn in outs["oracle"]:
        assert outs["oracle"][fn] == outs["b200"][fn], fn
    b200.clover_b200_invalidate_()
