#!/usr/bin/env python
"""SASS evidence for the seven production kernels of libclover_b200.so (cuobjdump -sass, sm_100a): per kernel the
instruction count, the counts of the Blackwell-specific tile-movement / synchronisation mnemonics (UTMALDG = the TMA
tile load cp.async.bulk.tensor, SYNCS.* = mbarrier arrive / try_wait, ACQBULK = async-proxy fence), the fp64 pipe mix
and -- as an excerpt -- the first TMA issue site and the mbarrier wait loop.
  python profiles/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cloverleaf_b200", "libclover_b200.so")
KERNELS = ["timestep_tma_kernel", "pdv_predict_eos_tma_kernel", "lagrange_correct_tma_kernel", "advec_cell_tma_kernel",
           "advec_cell_ymarch_tma_kernel", "advec_mom_ymarch_tma_kernel",
           "advec_mom_tma_kernel", "halo_exchange_kernel", "update_halo_kernel"]
MNEMONICS = ["UTMALDG", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "SYNCS.EXCH", "ACQBULK", "ATOMG", "LDS", "STS", "LDG", "STG", "DFMA",
             "DMUL", "DADD", "MUFU", "BAR.SYNC", "LDC", "MEMBAR", "ELECT"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    print("# cuobjdump -sass cloverleaf_b200/libclover_b200.so ; cubin architectures: %s" % ", ".join(arch))
    funcs = re.split(r"\n\s*Function : ", out)[1:]
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if not any(k in name for k in KERNELS):
            continue
        lines = [l for l in f.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
        ops = []
        for l in lines:
            toks = re.sub(r"/\*[0-9a-f]+\*/", "", l).split()
            toks = [t for t in toks if not t.startswith("@")]
            if toks:
                ops.append(toks[0].rstrip(";"))
        cnt = collections.Counter()
        for o in ops:
            for m in MNEMONICS:
                if o.startswith(m):
                    cnt[m] += 1
        short = re.sub(r"\(.*", "", demangled)
        print("\n## %s" % short)
        print("instructions %d ; " % len(ops) + " ".join("%s=%d" % (m, cnt[m]) for m in MNEMONICS if cnt[m]))
        # excerpt: the first TMA issue and the first mbarrier wait
        body = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in lines]
        for key, before, after in (("UTMALDG", 4, 2), ("SYNCS.PHASECHK", 1, 3)):
            idx = next((i for i, l in enumerate(body) if key in l), None)
            if idx is not None:
                print("  ... %s site:" % key)
                for l in body[max(0, idx - before):idx + after + 1]:
                    print("   " + l.strip())


if __name__ == "__main__":
    main()
