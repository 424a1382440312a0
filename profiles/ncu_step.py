"""Workload for ncu: N hydro steps of clover_bm16_short (3840^2) through the C-ABI, nothing else.
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
cloverleaf_b200.load_b200()
d = Driver(deck, cloverleaf_b200.LIB_B200, end_step=steps)
d.run()
print("ran", d.step, "steps; dt", d.dts()[-1])
