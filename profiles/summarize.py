"""Turn ncu artefacts brought back in gpurun_out/ into the small text summaries committed here.

  python profiles/summarize.py full  gpurun_out/prof.ncu-rep   > profiles/rNN_full.txt
  python profiles/summarize.py list  gpurun_out/launches.csv   > profiles/rNN_launches.txt
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_rd", "dram__bytes_read.sum"),
    ("dram_wr", "dram__bytes_write.sum"),
    ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("fp64_%", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("issue_%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("occ_%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("warp_inst", "smsp__inst_executed.sum"),
    ("l1_hit%", "l1tex__t_sector_hit_rate.pct"),
    ("l2_hit%", "lts__t_sector_hit_rate.pct"),
]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    col = {h: i for i, h in enumerate(H)}
    print("# ncu --set full --clock-control none ; one line per captured launch (units as reported by ncu)")
    print("kernel | " + " | ".join("%s[%s]" % (n, U[col[m]]) if m in col else n for n, m in METRICS))
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        vals = []
        for n, m in METRICS:
            v = r[col[m]] if m in col else "n/a"
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
            vals.append(v)
        print(name + " | " + " | ".join(vals))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r]
    h = next(i for i, r in enumerate(rows) if r[0] == "ID")
    H = rows[h]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) > vi:
            agg.setdefault(r[ki].split("(")[0].replace("void ", ""), []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    unit = rows[h + 1][ui]
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold cache: compare SHARES)")
    print("kernel | launches | mean[%s] | share of captured time" % unit)
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%s | %d | %.1f | %.3f" % (k, len(v), sum(v) / len(v), sum(v) / tot))
    print("total | %d | %.1f | 1.000" % (sum(len(v) for v in agg.values()), tot))


if __name__ == "__main__":
    {"full": full, "list": launches}[sys.argv[1]](sys.argv[2])
