"""Workload for ncu / compute-sanitizer: N hydro steps of clover_bm16_short (3840^2) through the C-ABI, nothing else.
  compute-sanitizer --tool racecheck python profiles/ncu_step.py 3 clover_bm_short.in 250 130
  ncu --set full --clock-control none --import-source on -k regex:'fused|advec' -s 16 -c 8 -o gpurun_out/prof \
      python profiles/ncu_step.py 4
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cloverleaf_b200
from cloverleaf_b200.driver import Driver

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
deck = sys.argv[2] if len(sys.argv) > 2 else "clover_bm16_short.in"
if len(sys.argv) > 4:  # mesh override for the slow tools:  ncu_step.py 3 clover_bm_short.in 250 130
    from cloverleaf_b200.driver import deck_text
    deck = deck_text(deck).replace("x_cells=960", "x_cells=%s" % sys.argv[3]).replace("y_cells=960", "y_cells=%s" % sys.argv[4])
cloverleaf_b200.load_b200()
d = Driver(deck, cloverleaf_b200.LIB_B200, end_step=steps)
d.run()
print("ran", d.step, "steps; dt", d.dts()[-1])
