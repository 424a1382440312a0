"""Python (ctypes) mirror of the kernel C-ABI declared in include/clover_b200.h.

The same table drives every backend exporting the reference's `*_kernel_c_` symbols
(CloverLeaf_ref/kernels/*_kernel_c.c): the CUDA library, the oracle port and oracle/_ref.
Argument order is the reference's; every argument is passed by reference.

Type codes:  i/d  = int / double scalar (by reference)
             C V X Y = 2-D field, Fortran shape cell (nx+4,ny+4) / vertex (nx+5,ny+5) /
                       x-face (nx+5,ny+4) / y-face (nx+4,ny+5), lower bound (-1,-1), stored here as
                       numpy [k+1, j+1] (C order == Fortran column-major with j fastest)
             W     = work array (vertex shape)
             cx vx cy vy = 1-D geometry arrays of length nx+4 / nx+5 / ny+4 / ny+5
             i4 i15 = int arrays;  s = per-state double array;  si = per-state int array
             buf   = 1-D message buffer
"""
import ctypes

import numpy as np

X4 = [("x_min", "i"), ("x_max", "i"), ("y_min", "i"), ("y_max", "i")]

KERNELS = {
    "ideal_gas_kernel_c_": X4 + [("density", "C"), ("energy", "C"), ("pressure", "C"), ("soundspeed", "C")],
    "viscosity_kernel_c_": X4 + [("celldx", "cx"), ("celldy", "cy"), ("density0", "C"), ("pressure", "C"),
                                 ("viscosity", "C"), ("xvel0", "V"), ("yvel0", "V")],
    "calc_dt_kernel_c_": X4 + [("g_small", "d"), ("g_big", "d"), ("dtmin", "d"), ("dtc_safe", "d"),
                               ("dtu_safe", "d"), ("dtv_safe", "d"), ("dtdiv_safe", "d"), ("xarea", "X"),
                               ("yarea", "Y"), ("cellx", "cx"), ("celly", "cy"), ("celldx", "cx"),
                               ("celldy", "cy"), ("volume", "C"), ("density0", "C"), ("energy0", "C"),
                               ("pressure", "C"), ("viscosity", "C"), ("soundspeed", "C"), ("xvel0", "V"),
                               ("yvel0", "V"), ("dt_min", "W"), ("dt_min_val", "d"), ("dtl_control", "i"),
                               ("xl_pos", "d"), ("yl_pos", "d"), ("jldt", "i"), ("kldt", "i"), ("small", "i")],
    "pdv_kernel_c_": [("prdct", "i")] + X4 + [("dt", "d"), ("xarea", "X"), ("yarea", "Y"), ("volume", "C"),
                                              ("density0", "C"), ("density1", "C"), ("energy0", "C"),
                                              ("energy1", "C"), ("pressure", "C"), ("viscosity", "C"),
                                              ("xvel0", "V"), ("xvel1", "V"), ("yvel0", "V"), ("yvel1", "V"),
                                              ("volume_change", "W")],
    "revert_kernel_c_": X4 + [("density0", "C"), ("density1", "C"), ("energy0", "C"), ("energy1", "C")],
    "accelerate_kernel_c_": X4 + [("dt", "d"), ("xarea", "X"), ("yarea", "Y"), ("volume", "C"),
                                  ("density0", "C"), ("pressure", "C"), ("viscosity", "C"), ("xvel0", "V"),
                                  ("yvel0", "V"), ("xvel1", "V"), ("yvel1", "V")],
    "flux_calc_kernel_c_": X4 + [("dt", "d"), ("xarea", "X"), ("yarea", "Y"), ("xvel0", "V"), ("yvel0", "V"),
                                 ("xvel1", "V"), ("yvel1", "V"), ("vol_flux_x", "X"), ("vol_flux_y", "Y")],
    "advec_cell_kernel_c_": X4 + [("dir", "i"), ("sweep_number", "i"), ("vertexdx", "vx"), ("vertexdy", "vy"),
                                  ("volume", "C"), ("density1", "C"), ("energy1", "C"), ("mass_flux_x", "X"),
                                  ("vol_flux_x", "X"), ("mass_flux_y", "Y"), ("vol_flux_y", "Y")] +
                            [("work%d" % i, "W") for i in range(1, 8)],
    "advec_mom_kernel_c_": X4 + [("vel1", "V"), ("mass_flux_x", "X"), ("vol_flux_x", "X"), ("mass_flux_y", "Y"),
                                 ("vol_flux_y", "Y"), ("volume", "C"), ("density1", "C")] +
                           [("work%d" % i, "W") for i in range(1, 7)] +
                           [("celldx", "cx"), ("celldy", "cy"), ("which_vel", "i"), ("sweep_number", "i"),
                            ("direction", "i")],
    "reset_field_kernel_c_": X4 + [("density0", "C"), ("density1", "C"), ("energy0", "C"), ("energy1", "C"),
                                   ("xvel0", "V"), ("xvel1", "V"), ("yvel0", "V"), ("yvel1", "V")],
    "update_halo_kernel_c_": X4 + [("chunk_neighbours", "i4"), ("tile_neighbours", "i4"), ("density0", "C"),
                                   ("energy0", "C"), ("pressure", "C"), ("viscosity", "C"), ("soundspeed", "C"),
                                   ("density1", "C"), ("energy1", "C"), ("xvel0", "V"), ("yvel0", "V"),
                                   ("xvel1", "V"), ("yvel1", "V"), ("vol_flux_x", "X"), ("vol_flux_y", "Y"),
                                   ("mass_flux_x", "X"), ("mass_flux_y", "Y"), ("fields", "i15"), ("depth", "i")],
    "field_summary_kernel_c_": X4 + [("volume", "C"), ("density0", "C"), ("energy0", "C"), ("pressure", "C"),
                                     ("xvel0", "V"), ("yvel0", "V"), ("vol", "d"), ("mass", "d"), ("ie", "d"),
                                     ("ke", "d"), ("press", "d")],
    "initialise_chunk_kernel_c_": X4 + [("min_x", "d"), ("min_y", "d"), ("dx", "d"), ("dy", "d"),
                                        ("vertexx", "vx"), ("vertexdx", "vx"), ("vertexy", "vy"),
                                        ("vertexdy", "vy"), ("cellx", "cx"), ("celldx", "cx"), ("celly", "cy"),
                                        ("celldy", "cy"), ("volume", "C"), ("xarea", "X"), ("yarea", "Y")],
    "generate_chunk_kernel_c_": X4 + [("vertexx", "vx"), ("vertexy", "vy"), ("cellx", "cx"), ("celly", "cy"),
                                      ("density0", "C"), ("energy0", "C"), ("xvel0", "V"), ("yvel0", "V"),
                                      ("number_of_states", "i"), ("state_density", "s"), ("state_energy", "s"),
                                      ("state_xvel", "s"), ("state_yvel", "s"), ("state_xmin", "s"),
                                      ("state_xmax", "s"), ("state_ymin", "s"), ("state_ymax", "s"),
                                      ("state_radius", "s"), ("state_geometry", "si"), ("g_rect", "i"),
                                      ("g_circ", "i"), ("g_point", "i")],
}
_PACK = X4 + [("field", "F"), ("buffer", "buf"), ("cell_data", "i"), ("vertex_data", "i"), ("x_face_data", "i"),
              ("y_face_data", "i"), ("depth", "i"), ("field_type", "i"), ("buffer_offset", "i")]
for _face in ("left", "right", "top", "bottom"):
    KERNELS["clover_pack_message_%s_c_" % _face] = _PACK
    KERNELS["clover_unpack_message_%s_c_" % _face] = _PACK

KERNEL_SYMBOLS = list(KERNELS)

EXTENSION_SYMBOLS = [
    "timer_c_", "clover_b200_init_", "clover_b200_finalize_", "clover_b200_set_resident_",
    "clover_b200_invalidate_", "clover_b200_forget_", "clover_b200_upload_", "clover_b200_download_", "clover_b200_sync_to_host_",
    "clover_b200_device_synchronize_", "clover_b200_register_chunk_", "clover_b200_comm_get_unique_id_",
    "clover_b200_comm_init_", "clover_b200_exchange_", "clover_b200_min_", "clover_b200_sum_",
    "clover_b200_launch_count_", "clover_b200_profile_", "clover_b200_profile_get_",
    "clover_b200_profile_reset_", "clover_b200_copy_bytes_", "clover_b200_halo_bytes_", "clover_b200_event_record_",
    "clover_b200_event_elapsed_ms_", "clover_b200_pin_", "clover_b200_unpin_", "clover_b200_selftest_math_",
    "clover_b200_set_fusion_",
    "clover_b200_set_tma_", "clover_b200_trace_", "clover_b200_trace_dump_", "clover_b200_transport_",
]

CELL_DATA, VERTEX_DATA, X_FACE_DATA, Y_FACE_DATA = 1, 2, 3, 4  # data.f90:68-71


def shape(code, nx, ny):
    """numpy shape [rows(k), cols(j)] of a 2-D field of the given type code."""
    return {"C": (ny + 4, nx + 4), "V": (ny + 5, nx + 5), "W": (ny + 5, nx + 5),
            "X": (ny + 4, nx + 5), "Y": (ny + 5, nx + 4)}[code]


def length(code, nx, ny):
    return {"cx": nx + 4, "vx": nx + 5, "cy": ny + 4, "vy": ny + 5}[code]


def call(lib, name, **kw):
    """Call kernel `name` of ctypes library `lib`.  Scalars may be Python numbers (in) or
    1-element numpy arrays (in/out); arrays must be C-contiguous float64 / int32 numpy arrays."""
    spec = KERNELS[name]
    args, keep = [], []
    for arg, code in spec:
        if arg not in kw:
            raise TypeError("%s: missing argument %s" % (name, arg))
        v = kw[arg]
        if code in ("i", "d"):
            if isinstance(v, np.ndarray):
                want = np.int32 if code == "i" else np.float64
                assert v.dtype == want and v.size == 1, (name, arg)
                args.append(ctypes.c_void_p(v.ctypes.data))
            else:
                c = ctypes.c_int(int(v)) if code == "i" else ctypes.c_double(float(v))
                keep.append(c)
                args.append(ctypes.byref(c))
        else:
            want = np.int32 if code in ("i4", "i15", "si") else np.float64
            assert isinstance(v, np.ndarray) and v.dtype == want and v.flags["C_CONTIGUOUS"], (name, arg)
            args.append(ctypes.c_void_p(v.ctypes.data))
    fn = getattr(lib, name)
    fn.restype = None
    fn(*args)
