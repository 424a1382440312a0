"""integration/cloverleaf_ref_b200.patch -- the Fortran side of the drop-in as an artefact a maintainer
can apply to CloverLeaf_ref.  No Fortran compiler or MPI exists in this image, so it is checked as far as that allows:

  * it applies cleanly to the reference tree (where /root/reference exists: this container);
  * every `CALL clover_b200_*` it introduces is declared in include/clover_b200_kernels.f90 with the same number of
    arguments, is named in the USE ... ONLY list of the file that calls it, and is exported by libclover_b200.so;
  * it touches the routines the survey names: clover_init_comms, clover_finalize, clover_exchange, clover_sum,
    clover_min (clover.f90), the chunk registration in start.f90, the D2H hook in visit.f90, the five sums of
    field_summary.f90, and adds a link target to the Makefile.
"""
import ctypes
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

PATCH = os.path.join(ROOT, "integration", "cloverleaf_ref_b200.patch")
REF = "/root/reference/CloverLeaf_ref"
F90 = os.path.join(ROOT, "include", "clover_b200_kernels.f90")


def _added_by_file():
    out, cur = {}, None
    for line in open(PATCH).read().splitlines():
        if line.startswith("+++ "):
            cur = line[4:].split("/", 1)[1].strip()
            out[cur] = []
        elif line.startswith("+") and not line.startswith("+++") and cur:
            out[cur].append(line[1:])
    return out


def _join_continuations(lines):
    text, buf = [], ""
    for l in lines:
        s = l.split("!")[0].rstrip() if not l.lstrip().startswith("!") else ""
        if not s.strip():
            continue
        if s.rstrip().endswith("&"):
            buf += s.rstrip()[:-1] + " "
        else:
            text.append(buf + s)
            buf = ""
    return text


def _split_args(argstr):
    args, depth, cur = [], 0, ""
    for ch in argstr:
        if ch == "(":
            depth += 1
        if ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            args.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        args.append(cur.strip())
    return args


def test_patch_touches_the_named_routines():
    added = _added_by_file()
    assert set(added) == {"Makefile", "clover.f90", "field_summary.f90", "start.f90", "visit.f90"}
    clover = "\n".join(added["clover.f90"])
    for name in ("clover_b200_init", "clover_b200_comm_get_unique_id", "clover_b200_comm_init", "clover_b200_finalize",
                 "clover_b200_exchange", "clover_b200_sum", "clover_b200_min"):
        assert "CALL %s(" % name in clover, name
    assert "CALL clover_b200_register_chunk(" in "\n".join(added["start.f90"])
    assert "CALL clover_b200_sync_to_host(" in "\n".join(added["visit.f90"])
    assert "CALL clover_b200_sum(b200_sums,5)" in "\n".join(added["field_summary.f90"])
    mk = "\n".join(added["Makefile"])
    assert "clover_leaf_b200:" in mk and "-lclover_b200" in mk and "clover_b200_kernels.f90" in mk
    assert "_kernel_c.o" not in mk  # the C kernel objects are what the library replaces


def test_every_introduced_call_is_declared_imported_and_exported():
    import cloverleaf_b200
    module = open(F90).read()
    lib = ctypes.CDLL(cloverleaf_b200.LIB_B200)  # loading needs no GPU
    decl = {}
    for m in re.finditer(r"SUBROUTINE (clover_b200_\w+)\((.*?)\)\s*(?:&\s*\n\s*)?BIND\(C, NAME='(\w+)'\)", module, re.S):
        args = [a for a in m.group(2).replace("&", " ").replace("\n", " ").split(",") if a.strip()]
        decl[m.group(1)] = (len(args), m.group(3))
    seen = 0
    for fname, lines in _added_by_file().items():
        if not fname.endswith(".f90"):
            continue
        stmts = _join_continuations(lines)
        only = " ".join(s for s in stmts if "USE clover_b200_kernels" in s)
        for s in stmts:
            for m in re.finditer(r"CALL (clover_b200_\w+)\((.*)\)\s*$", s):
                name, nargs = m.group(1), len(_split_args(m.group(2)))
                assert name in decl, "%s: %s is not in the ISO_C_BINDING module" % (fname, name)
                assert decl[name][0] == nargs, "%s: %s called with %d arguments, declared with %d" % (
                    fname, name, nargs, decl[name][0])
                assert re.search(r"\b%s\b" % name, only), "%s: %s missing from USE clover_b200_kernels, ONLY:" % (fname, name)
                getattr(lib, decl[name][1])
                seen += 1
    assert seen >= 9


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists only in the build container")
def test_patch_applies_cleanly_to_the_reference(tmp_path):
    if not shutil.which("patch"):
        pytest.skip("no patch(1)")
    dst = tmp_path / "CloverLeaf_ref"
    shutil.copytree(REF, dst)
    for root, dirs, files in os.walk(dst):
        for n in dirs + files:
            os.chmod(os.path.join(root, n), 0o755 if n in dirs else 0o644)
    r = subprocess.run(["patch", "-p1", "--dry-run", "-i", PATCH], cwd=dst, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FAILED" not in r.stdout and "fuzz" not in r.stdout, r.stdout
    r = subprocess.run(["patch", "-p1", "-i", PATCH], cwd=dst, capture_output=True, text=True)
    assert r.returncode == 0
    # the hooks sit inside the routines they replace
    src = open(dst / "clover.f90").read()
    ex = src[src.index("SUBROUTINE clover_exchange"):src.index("END SUBROUTINE clover_exchange")]
    assert "CALL clover_b200_exchange(fields,depth)" in ex and ex.index("clover_b200_exchange") < ex.index("request=0")
    mn = src[src.index("SUBROUTINE clover_min"):src.index("END SUBROUTINE clover_min")]
    assert "CALL clover_b200_min(value)" in mn
    ini = src[src.index("SUBROUTINE clover_init_comms"):src.index("END SUBROUTINE clover_init_comms")]
    assert ini.index("MPI_INIT") < ini.index("clover_b200_init") < ini.index("clover_b200_comm_init")
