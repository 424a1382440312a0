"""Multi-process tests of the N > 1 path.

CPU (gloo, world_size 2 and 4; runs everywhere): one chunk per process, halo messages through the ABI
pack/unpack kernels + torch.distributed point-to-point (what MPI_ISEND/IRECV are to the Fortran
driver, clover.f90:348-500), dt by all_reduce(MIN), summaries by all_reduce(SUM).  Checked against
the single-process run of the same deck: dt bit-identical at every step.

GPU (nccl; needs >= 2 devices, marked gpu): the same through libclover_b200.so's device pack +
ncclSend/ncclRecv exchange (clover_b200_exchange_) and ncclAllReduce reductions.
"""
import ctypes
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ORACLE_PORT, ROOT

WORKER = r'''
import ctypes, json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["CLV_ROOT"])
from cloverleaf_b200.driver import Driver, deck_text
import cloverleaf_b200

backend = os.environ["CLV_BACKEND"]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
deck = deck_text("clover_bm_short.in").replace("x_cells=960", "x_cells=%s" % os.environ["CLV_NX"]).replace(
    "y_cells=960", "y_cells=%s" % os.environ["CLV_NY"])
steps = int(os.environ["CLV_STEPS"])
if backend == "gloo":
    dist.init_process_group("gloo")
    d = Driver(deck, os.environ["CLV_LIB"], nchunks=world, rank=rank, comm_mode=2, end_step=steps)

    def sendrecv(peer, snd, rcv, count):
        s = torch.from_numpy(np.ctypeslib.as_array(snd, shape=(count,)).copy())
        r = torch.empty(count, dtype=torch.float64)
        reqs = [dist.isend(s, peer), dist.irecv(r, peer)]
        for q in reqs:
            q.wait()
        np.ctypeslib.as_array(rcv, shape=(count,))[:] = r.numpy()

    def allreduce(values, n, op):
        t = torch.from_numpy(np.ctypeslib.as_array(values, shape=(n,)).copy())
        dist.all_reduce(t, op=dist.ReduceOp.MIN if op == 0 else dist.ReduceOp.SUM)
        np.ctypeslib.as_array(values, shape=(n,))[:] = t.numpy()

    d.set_comm_callbacks(sendrecv, allreduce)
else:
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = cloverleaf_b200.load_b200()
    dev = ctypes.c_int(local)
    lib.clover_b200_init_(ctypes.byref(dev))
    idbuf = ctypes.create_string_buffer(128)
    if rank == 0:
        lib.clover_b200_comm_get_unique_id_(idbuf)
    t = torch.tensor(list(idbuf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    idbuf = ctypes.create_string_buffer(bytes(t.cpu().tolist()), 128)
    nr, rk = ctypes.c_int(world), ctypes.c_int(rank)
    lib.clover_b200_comm_init_(ctypes.byref(nr), ctypes.byref(rk), idbuf)
    d = Driver(deck, cloverleaf_b200.LIB_B200, nchunks=world, rank=rank, comm_mode=1, end_step=steps)
die_at = int(os.environ.get("CLV_DIE_AT", "0"))
if die_at:
    d.run(die_at)
    if rank == 1:
        os._exit(0)  # a rank that dies mid-run: its neighbours must abort with a diagnostic, not hang the GPU
d.run()
out = dict(rank=rank, dt=d.dts().tolist(), summaries=d.summaries(), chunk=d.chunk_info(0))
with open(os.path.join(os.environ["CLV_OUT"], "rank%d.json" % rank), "w") as f:
    json.dump(out, f)
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_world(tmp_path, world, backend, lib, nx, ny, steps, extra_env=None):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, **(extra_env or {}))
    env = dict(env, CLV_ROOT=ROOT, CLV_BACKEND=backend, CLV_LIB=lib, CLV_NX=str(nx), CLV_NY=str(ny),
               CLV_STEPS=str(steps), CLV_OUT=str(tmp_path), OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return [json.load(open(tmp_path / ("rank%d.json" % i))) for i in range(world)]


def _single(lib, nx, ny, steps):
    from cloverleaf_b200.driver import Driver, deck_text
    deck = deck_text("clover_bm_short.in").replace("x_cells=960", "x_cells=%d" % nx).replace(
        "y_cells=960", "y_cells=%d" % ny)
    d = Driver(deck, lib, end_step=steps)
    d.run()
    return d.dts().tolist(), d.summaries()


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_ranks_match_single_process(tmp_path, world):
    nx, ny, steps = 48, 40, 25
    ranks = _run_world(tmp_path, world, "gloo", ORACLE_PORT, nx, ny, steps)
    dt1, s1 = _single(ORACLE_PORT, nx, ny, steps)
    for r in ranks:
        assert r["dt"] == dt1, "rank %d: dt differs from the single-process run" % r["rank"]
    # all_reduce(SUM) leaves the global sums on every rank (the reference reduces to rank 0)
    for a, b in zip(ranks[0]["summaries"], s1):
        for k in ("volume", "mass", "ie", "ke", "pressure"):
            assert abs(a[k] - b[k]) <= 1e-12 * max(abs(b[k]), 1e-300)
    assert sum(r["chunk"]["x_max"] * r["chunk"]["y_max"] for r in ranks) == nx * ny


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20)
        return out.stdout.count("GPU ") if out.returncode == 0 else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_ranks_match_single_gpu(tmp_path, world):
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import cloverleaf_b200
    nx, ny, steps = 256, 192, 30
    ranks = _run_world(tmp_path, world, "nccl", cloverleaf_b200.LIB_B200, nx, ny, steps)
    dt1, s1 = _single(ORACLE_PORT, nx, ny, steps)
    for r in ranks:
        assert r["dt"] == dt1, "rank %d: dt differs from the oracle's single-chunk run" % r["rank"]
    for a, b in zip(ranks[0]["summaries"], s1):
        for k in ("volume", "mass", "ie", "ke", "pressure"):
            assert abs(a[k] - b[k]) <= 1e-10 * max(abs(b[k]), 1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("world,nx,ny", [(2, 250, 130), (4, 250, 130), (4, 61, 37)])
def test_ragged_chunks_match_single_gpu(tmp_path, world, nx, ny):
    """Odd mesh sizes: 2 x 125x130, 4 x 125x65 and 4 chunks of 31/30 x 19/18 cells (clover_decompose gives the
    remainder to the low-index chunks, clover.f90:150-170; unequal neighbours, tiles that are mostly rim, strips
    shorter than a copy segment) through the peer-memory exchange: dt bit-identical to the oracle's single-chunk run."""
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import cloverleaf_b200
    steps = 25
    ranks = _run_world(tmp_path, world, "nccl", cloverleaf_b200.LIB_B200, nx, ny, steps)
    dt1, s1 = _single(ORACLE_PORT, nx, ny, steps)
    for r in ranks:
        assert r["dt"] == dt1, "rank %d: dt differs from the oracle's single-chunk run" % r["rank"]
    for a, b in zip(ranks[0]["summaries"], s1):
        for k in ("volume", "mass", "ie", "ke", "pressure"):
            assert abs(a[k] - b[k]) <= 1e-10 * max(abs(b[k]), 1e-300)
    assert sum(r["chunk"]["x_max"] * r["chunk"]["y_max"] for r in ranks) == nx * ny


@pytest.mark.gpu
def test_nccl_transport_matches_single_gpu(tmp_path):
    """The fallback transport (ncclSend/ncclRecv + ncclAllReduce, CLOVER_B200_P2P=0) against the same oracle run."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    import cloverleaf_b200
    nx, ny, steps = 256, 192, 30
    ranks = _run_world(tmp_path, 2, "nccl", cloverleaf_b200.LIB_B200, nx, ny, steps,
                       extra_env={"CLOVER_B200_P2P": "0"})
    dt1, _ = _single(ORACLE_PORT, nx, ny, steps)
    for r in ranks:
        assert r["dt"] == dt1, "rank %d: dt differs from the oracle's single-chunk run" % r["rank"]


@pytest.mark.gpu
def test_dead_rank_aborts_the_others(tmp_path):
    """A rank that disappears mid-run: the surviving rank's exchange kernel gives up after the spin time-out, leaves
    an error record, traps, and the host aborts with a diagnostic -- instead of spinning on a flag for ever with a
    GPU that no longer answers (the peer-memory protocol has no other failure channel)."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    import time
    import cloverleaf_b200
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, CLV_ROOT=ROOT, CLV_BACKEND="nccl", CLV_LIB=cloverleaf_b200.LIB_B200, CLV_NX="256",
               CLV_NY="192", CLV_STEPS="30", CLV_OUT=str(tmp_path), OMP_NUM_THREADS="1", CLV_DIE_AT="7",
               CLOVER_B200_SPIN_TIMEOUT_MS="3000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    t0 = time.time()
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    took = time.time() - t0
    assert r.returncode != 0
    assert "device-side time-out" in r.stderr and "halo exchange" in r.stderr, r.stderr[-3000:]
    assert not os.path.exists(tmp_path / "rank0.json")
    assert took < 120, took
