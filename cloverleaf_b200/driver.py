"""ctypes binding of libclover_driver.so (the C++ restatement of the Fortran driver).

The driver is backend-agnostic: it dlopens any shared library exporting the
reference's `*_kernel_c_` symbols (CloverLeaf_ref/kernels/*_kernel_c.c).
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DRIVER = os.path.join(HERE, "libclover_driver.so")
DECK_DIR = os.path.join(HERE, "decks")

from .abi import KERNEL_SYMBOLS, EXTENSION_SYMBOLS  # noqa: F401  (re-exported)

_lib = None


def _driver_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_DRIVER):
            raise RuntimeError("libclover_driver.so is not built (python -m cloverleaf_b200.build)")
        L = ctypes.CDLL(LIB_DRIVER)
        vp, ci, cd, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_char_p
        L.clover_driver_create.restype = vp
        L.clover_driver_create.argtypes = [cs, cs, ci, ci, ci, cs]
        L.clover_driver_error.restype = cs
        L.clover_driver_error.argtypes = [vp]
        for name, res, args in [
            ("clover_driver_set_end_step", None, [vp, ci]),
            ("clover_driver_set_end_time", None, [vp, cd]),
            ("clover_driver_set_summary_frequency", None, [vp, ci]),
            ("clover_driver_set_visit", None, [vp, cs, ci]),
            ("clover_driver_start", ci, [vp]),
            ("clover_driver_run", ci, [vp, ci]),
            ("clover_driver_field_summary", None, [vp]),
            ("clover_driver_complete", ci, [vp]),
            ("clover_driver_step", ci, [vp]),
            ("clover_driver_time", cd, [vp]),
            ("clover_driver_dt", cd, [vp]),
            ("clover_driver_wall", cd, [vp]),
            ("clover_driver_num_chunks", ci, [vp]),
            ("clover_driver_grid", None, [vp, vp]),
            ("clover_driver_num_steps", ci, [vp]),
            ("clover_driver_get_steps", None, [vp, vp]),
            ("clover_driver_num_summaries", ci, [vp]),
            ("clover_driver_get_summaries", None, [vp, vp]),
            ("clover_driver_chunk_info", None, [vp, ci, vp]),
            ("clover_driver_field", vp, [vp, ci, cs]),
            ("clover_driver_sync_to_host", None, [vp]),
            ("clover_driver_destroy", None, [vp]),
            ("clover_driver_set_comm_callbacks", None, [vp, vp, vp]),
        ]:
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def deck_text(name_or_text):
    """Return deck text: either literal text containing '*clover' or a name in decks/."""
    if "*clover" in name_or_text:
        return name_or_text
    path = name_or_text
    if not os.path.exists(path):
        path = os.path.join(DECK_DIR, name_or_text)
    with open(path) as f:
        return f.read()


SUMMARY_COLS = ["step", "time", "volume", "mass", "density", "pressure", "ie", "ke", "total"]
FIELD_SHAPES = {  # (extra x, extra y) over (nx+4, ny+4); build_field.f90:33-94
    "density0": (0, 0), "density1": (0, 0), "energy0": (0, 0), "energy1": (0, 0),
    "pressure": (0, 0), "viscosity": (0, 0), "soundspeed": (0, 0), "volume": (0, 0),
    "xvel0": (1, 1), "xvel1": (1, 1), "yvel0": (1, 1), "yvel1": (1, 1),
    "vol_flux_x": (1, 0), "mass_flux_x": (1, 0), "xarea": (1, 0),
    "vol_flux_y": (0, 1), "mass_flux_y": (0, 1), "yarea": (0, 1),
}


class Driver:
    """One CloverLeaf run: deck + kernel backend (.so path) [+ chunk decomposition]."""

    def __init__(self, deck, backend_so, nchunks=1, rank=0, comm_mode=0, log_path=None,
                 end_step=None, summary_frequency=None):
        L = _driver_lib()
        self._L = L
        self._h = L.clover_driver_create(deck_text(deck).encode(), str(backend_so).encode(),
                                         int(nchunks), int(rank), int(comm_mode),
                                         (log_path or "").encode())
        self._check()
        if end_step is not None:
            L.clover_driver_set_end_step(self._h, int(end_step))
        if summary_frequency is not None:
            L.clover_driver_set_summary_frequency(self._h, int(summary_frequency))
        self.started = False

    SENDRECV = ctypes.CFUNCTYPE(None, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                ctypes.POINTER(ctypes.c_double), ctypes.c_int)
    ALLREDUCE = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int)

    def set_visit(self, directory, frequency=0):
        """visit.f90 output (clover.visit + one ASCII VTK file per chunk and dump) into `directory`, every
        `frequency` steps (0: the deck's visit_frequency).  Call before start()."""
        self._L.clover_driver_set_visit(self._h, str(directory).encode(), int(frequency))

    def set_comm_callbacks(self, sendrecv, allreduce):
        """comm_mode=2: sendrecv(peer, snd, rcv, count) and allreduce(values, n, op[0 min,1 sum]) are
        Python callables working on ctypes double pointers (what MPI is to the Fortran driver)."""
        self._cb = (self.SENDRECV(sendrecv), self.ALLREDUCE(allreduce))  # keep alive
        self._L.clover_driver_set_comm_callbacks(self._h, ctypes.cast(self._cb[0], ctypes.c_void_p),
                                                 ctypes.cast(self._cb[1], ctypes.c_void_p))

    def _check(self):
        err = self._L.clover_driver_error(self._h).decode()
        if err:
            raise RuntimeError("clover_driver: " + err)

    def start(self):
        self._L.clover_driver_start(self._h)
        self._check()
        self.started = True
        return self

    def run(self, nsteps=1 << 30):
        if not self.started:
            self.start()
        n = self._L.clover_driver_run(self._h, int(nsteps))
        self._check()
        return n

    def field_summary(self):
        self._L.clover_driver_field_summary(self._h)
        return self.summaries()[-1]

    @property
    def complete(self):
        return bool(self._L.clover_driver_complete(self._h))

    @property
    def step(self):
        return self._L.clover_driver_step(self._h)

    @property
    def time(self):
        return self._L.clover_driver_time(self._h)

    @property
    def wall(self):
        return self._L.clover_driver_wall(self._h)

    def grid(self):
        a = (ctypes.c_int * 4)()
        self._L.clover_driver_grid(self._h, a)
        return dict(x_cells=a[0], y_cells=a[1], chunk_x=a[2], chunk_y=a[3])

    def steps(self):
        n = self._L.clover_driver_num_steps(self._h)
        a = np.zeros((n, 3))
        if n:
            self._L.clover_driver_get_steps(self._h, a.ctypes.data)
        return a

    def dts(self):
        return self.steps()[:, 2].copy()

    def summaries(self):
        n = self._L.clover_driver_num_summaries(self._h)
        a = np.zeros((n, 9))
        if n:
            self._L.clover_driver_get_summaries(self._h, a.ctypes.data)
        return [dict(zip(SUMMARY_COLS, row.tolist())) for row in a]

    def num_local_chunks(self):
        return self._L.clover_driver_num_chunks(self._h)

    def chunk_info(self, idx=0):
        a = (ctypes.c_int * 11)()
        self._L.clover_driver_chunk_info(self._h, idx, a)
        keys = ["id", "left", "right", "bottom", "top", "x_max", "y_max", "nb_left", "nb_right",
                "nb_bottom", "nb_top"]
        return dict(zip(keys, list(a)))

    def sync_to_host(self):
        self._L.clover_driver_sync_to_host(self._h)

    def field(self, name, idx=0):
        """Copy of a local chunk's field as a (ny+4[+1], nx+4[+1]) array, [k+1, j+1] indexing."""
        info = self.chunk_info(idx)
        ex, ey = FIELD_SHAPES[name]
        nx, ny = info["x_max"] + 4 + ex, info["y_max"] + 4 + ey
        p = self._L.clover_driver_field(self._h, idx, name.encode())
        buf = (ctypes.c_double * (nx * ny)).from_address(p)
        return np.frombuffer(buf, dtype=np.float64).reshape(ny, nx).copy()

    def global_field(self, name):
        """Assemble interior values of all local chunks (comm_mode 0) into the global mesh."""
        g = self.grid()
        ex, ey = FIELD_SHAPES[name]
        out = np.zeros((g["y_cells"] + ey, g["x_cells"] + ex))
        for i in range(self.num_local_chunks()):
            info = self.chunk_info(i)
            a = self.field(name, i)
            nx, ny = info["x_max"], info["y_max"]
            out[info["bottom"] - 1:info["bottom"] - 1 + ny + ey,
                info["left"] - 1:info["left"] - 1 + nx + ex] = a[2:2 + ny + ey, 2:2 + nx + ex]
        return out

    def close(self):
        if self._h:
            self._L.clover_driver_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
