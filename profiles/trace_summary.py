#!/usr/bin/env python
"""Summarise an in-situ launch timeline (bench.py --trace PREFIX -> PREFIX.rankN.csv, see clover_b200_trace_).

For every kernel name: launches, mean duration (last CTA end - first CTA start), mean gap to the END of the previous
launch (negative = it started before the previous launch had finished: PDL overlap), mean time its CTAs spent waiting
for the preceding halo kernel, and the share of the traced span.  usage: trace_summary.py FILE.csv [...]"""
import csv
import sys
from collections import OrderedDict


def summarise(path):
    rows = list(csv.DictReader(open(path)))
    rows = [r for r in rows if r["start_ns"] and r["end_ns"]]
    for r in rows:
        for k in ("start_ns", "end_ns", "wait_begin_ns", "wait_end_ns"):
            r[k] = float(r[k]) if r[k] else None
    span = max(r["end_ns"] for r in rows) - min(r["start_ns"] for r in rows)
    by = OrderedDict()
    prev_end = None
    for r in rows:
        s = by.setdefault(r["name"], dict(n=0, dur=0.0, gap=0.0, ngap=0, wait=0.0, work=0.0))
        s["n"] += 1
        s["dur"] += r["end_ns"] - r["start_ns"]
        if prev_end is not None:
            s["gap"] += r["start_ns"] - prev_end
            s["ngap"] += 1
        if r["wait_begin_ns"] is not None and r["wait_end_ns"] is not None:
            s["wait"] += max(0.0, r["wait_end_ns"] - r["wait_begin_ns"])
        # time from "previous launch finished" (or own start, whichever is later) to own end = what this launch adds
        s["work"] += r["end_ns"] - max(r["start_ns"], prev_end if prev_end is not None else r["start_ns"])
        prev_end = max(prev_end, r["end_ns"]) if prev_end is not None else r["end_ns"]
    print("%s: %d launches, span %.3f ms" % (path, len(rows), span / 1e6))
    print("%-24s %6s %10s %10s %10s %10s %7s" % ("kernel", "n", "dur us", "gap us", "wait us", "adds us", "share"))
    for name, s in by.items():
        print("%-24s %6d %10.2f %10.2f %10.2f %10.2f %7.3f" % (
            name, s["n"], s["dur"] / s["n"] / 1e3, (s["gap"] / s["ngap"] / 1e3) if s["ngap"] else 0.0,
            s["wait"] / s["n"] / 1e3, s["work"] / s["n"] / 1e3, s["work"] / span))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        summarise(p)
