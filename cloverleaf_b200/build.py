"""Build recipes for the native pieces (all built IN-TREE so they travel with gpurun).

  libclover_b200.so    hand-written sm_100a CUDA kernels + the C-ABI (csrc/*.cu)
  libclover_driver.so  host driver: C++ restatement of the Fortran call sequence
  oracle/              test infrastructure: C port + (when /root/reference exists) oracle/_ref
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_B200 = os.path.join(HERE, "libclover_b200.so")
LIB_DRIVER = os.path.join(HERE, "libclover_driver.so")
ORACLE_DIR = os.path.join(ROOT, "oracle")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # fp64 parity: no FMA contraction, IEEE div/sqrt -> bit-identical to the
    # reference C kernels built with -ffp-contract=off (see DESIGN.md "Numerics").
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-shared",
    "-I" + os.path.join(ROOT, "include"),
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def cuda_sources():
    out = []
    for d, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h")):
                out.append(os.path.join(d, f))
    out.append(os.path.join(ROOT, "include", "clover_b200.h"))
    return out


def build_b200(force=False, verbose=False):
    srcs = cuda_sources()
    if not force and not _newer(LIB_B200, srcs):
        return LIB_B200
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cus = [s for s in srcs if s.endswith(".cu")]
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("CLV_NVCC_EXTRA", "").split() + ["--threads", "4", "-I/usr/include"] + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_B200] + sorted(cus) + ["-ldl"]
    out = _run(cmd)
    if verbose:
        print(out)
    return LIB_B200


def build_driver(force=False):
    src = os.path.join(CSRC, "driver", "clover_driver.cpp")
    if force or _newer(LIB_DRIVER, [src]):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", LIB_DRIVER, src, "-ldl"])
    return LIB_DRIVER


def build_oracle():
    """Test infrastructure: the C port always; oracle/_ref only where /root/reference exists."""
    _run(["make", "-C", ORACLE_DIR, "all"])


def build_all(force=False, verbose=False):
    build_driver(force)
    build_b200(force, verbose)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", LIB_B200, LIB_DRIVER)
