"""Emit include/clover_b200_kernels.f90: the ISO_C_BINDING interface module for every symbol of the kernel ABI
(north_star: "a thin ISO_C_BINDING C-ABI that mirrors the existing *_kernel_c entry points").

The argument lists come from cloverleaf_b200/abi.py (the table the parity tests call the library through), so the
module, the header and the tests cannot drift apart.  No Fortran compiler exists in the build image; the module is
interface-only and is checked textually by tests/test_fortran_interface.py.

    python include/gen_fortran_interface.py > include/clover_b200_kernels.f90
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cloverleaf_b200 import abi  # noqa: E402

# reference wrapper that makes the call (CloverLeaf_ref/*.f90), for the comments
CALLERS = {
    "ideal_gas_kernel_c_": "ideal_gas.f90:64,74", "viscosity_kernel_c_": "viscosity.f90:57",
    "calc_dt_kernel_c_": "calc_dt.f90:83", "pdv_kernel_c_": "PdV.f90:89", "revert_kernel_c_": "revert.f90:53",
    "accelerate_kernel_c_": "accelerate.f90:64", "flux_calc_kernel_c_": "flux_calc.f90:62",
    "advec_cell_kernel_c_": "advec_cell_driver.f90:59", "advec_mom_kernel_c_": "advec_mom_driver.f90:85,108",
    "reset_field_kernel_c_": "reset_field.f90:63", "update_halo_kernel_c_": "update_halo.f90:86",
    "field_summary_kernel_c_": "field_summary.f90:91", "initialise_chunk_kernel_c_": "initialise_chunk.f90:61",
    "generate_chunk_kernel_c_": "generate_chunk.f90:77",
}
INT_ARRAYS = {"i4": "(4)", "i15": "(15)", "si": "(*)"}


def wrap(prefix, names, width=100):
    lines, cur = [], prefix
    for i, n in enumerate(names):
        piece = n + ("," if i + 1 < len(names) else "")
        if len(cur) + len(piece) > width:
            lines.append(cur + " &")
            cur = " " * 8
        cur += piece
    lines.append(cur)
    return lines


def block(sym, spec):
    name = sym[:-1]  # Fortran name without the trailing underscore of the C symbol
    args = [a for a, _ in spec]
    out = ["    ! called from " + CALLERS[sym]] if sym in CALLERS else []
    head = wrap("    SUBROUTINE %s(" % name, args)
    head[-1] += ") &"
    out += head
    out.append("        BIND(C, NAME='%s')" % sym)
    out.append("      IMPORT :: C_INT, C_DOUBLE")
    ints = [a for a, c in spec if c == "i"]
    dbls = [a for a, c in spec if c == "d"]
    iarr = [a + INT_ARRAYS[c] for a, c in spec if c in INT_ARRAYS]
    darr = [a + "(*)" for a, c in spec if c not in ("i", "d") and c not in INT_ARRAYS]
    for kind, names in (("INTEGER(C_INT) :: ", ints), ("REAL(C_DOUBLE) :: ", dbls), ("INTEGER(C_INT) :: ", iarr),
                        ("REAL(C_DOUBLE) :: ", darr)):
        if names:
            out += wrap("      " + kind, names)
    out.append("    END SUBROUTINE %s" % name)
    return out


EXTENSION = r'''
    ! ---- extension: what the reference ABI has no word for (include/clover_b200.h) ---------------------------------
    SUBROUTINE clover_b200_init(device) BIND(C, NAME='clover_b200_init_')
      IMPORT :: C_INT
      INTEGER(C_INT) :: device
    END SUBROUTINE clover_b200_init
    SUBROUTINE clover_b200_finalize() BIND(C, NAME='clover_b200_finalize_')
    END SUBROUTINE clover_b200_finalize
    SUBROUTINE clover_b200_set_resident(on) BIND(C, NAME='clover_b200_set_resident_')
      IMPORT :: C_INT
      INTEGER(C_INT) :: on
    END SUBROUTINE clover_b200_set_resident
    SUBROUTINE clover_b200_set_fusion(on) BIND(C, NAME='clover_b200_set_fusion_')
      IMPORT :: C_INT
      INTEGER(C_INT) :: on
    END SUBROUTINE clover_b200_set_fusion
    SUBROUTINE clover_b200_set_tma(on) BIND(C, NAME='clover_b200_set_tma_')
      IMPORT :: C_INT
      INTEGER(C_INT) :: on
    END SUBROUTINE clover_b200_set_tma
    SUBROUTINE clover_b200_invalidate() BIND(C, NAME='clover_b200_invalidate_')
    END SUBROUTINE clover_b200_invalidate
    SUBROUTINE clover_b200_forget(field) BIND(C, NAME='clover_b200_forget_')
      IMPORT :: C_DOUBLE
      REAL(C_DOUBLE) :: field(*)
    END SUBROUTINE clover_b200_forget
    SUBROUTINE clover_b200_upload(field) BIND(C, NAME='clover_b200_upload_')
      IMPORT :: C_DOUBLE
      REAL(C_DOUBLE) :: field(*)
    END SUBROUTINE clover_b200_upload
    SUBROUTINE clover_b200_download(field) BIND(C, NAME='clover_b200_download_')
      IMPORT :: C_DOUBLE
      REAL(C_DOUBLE) :: field(*)
    END SUBROUTINE clover_b200_download
    ! visit.f90:25-180 reads host arrays: call this first (fields(f) = 1 selects field id f, data.f90:51-66)
    SUBROUTINE clover_b200_sync_to_host(fields) BIND(C, NAME='clover_b200_sync_to_host_')
      IMPORT :: C_INT
      INTEGER(C_INT) :: fields(15)
    END SUBROUTINE clover_b200_sync_to_host
    SUBROUTINE clover_b200_device_synchronize() BIND(C, NAME='clover_b200_device_synchronize_')
    END SUBROUTINE clover_b200_device_synchronize
    ! start.f90:92-97, after build_field / clover_allocate_buffers: the chunk whose halos clover_b200_exchange moves
    SUBROUTINE clover_b200_register_chunk(x_min,x_max,y_min,y_max,chunk_neighbours,density0,density1,energy0, &
        energy1,pressure,viscosity,soundspeed,xvel0,xvel1,yvel0,yvel1,vol_flux_x,vol_flux_y,mass_flux_x, &
        mass_flux_y) BIND(C, NAME='clover_b200_register_chunk_')
      IMPORT :: C_INT, C_DOUBLE
      INTEGER(C_INT) :: x_min,x_max,y_min,y_max
      INTEGER(C_INT) :: chunk_neighbours(4)
      REAL(C_DOUBLE) :: density0(*),density1(*),energy0(*),energy1(*),pressure(*),viscosity(*),soundspeed(*), &
                        xvel0(*),xvel1(*),yvel0(*),yvel1(*),vol_flux_x(*),vol_flux_y(*),mass_flux_x(*),mass_flux_y(*)
    END SUBROUTINE clover_b200_register_chunk
    ! clover_init_comms (clover.f90:70-94): rank 0 gets the id, MPI_BCAST it, every rank calls comm_init
    SUBROUTINE clover_b200_comm_get_unique_id(id128) BIND(C, NAME='clover_b200_comm_get_unique_id_')
      IMPORT :: C_CHAR
      CHARACTER(KIND=C_CHAR) :: id128(128)
    END SUBROUTINE clover_b200_comm_get_unique_id
    SUBROUTINE clover_b200_comm_init(nranks, rank, id128) BIND(C, NAME='clover_b200_comm_init_')
      IMPORT :: C_INT, C_CHAR
      INTEGER(C_INT) :: nranks, rank
      CHARACTER(KIND=C_CHAR) :: id128(128)
    END SUBROUTINE clover_b200_comm_init
    ! replaces the body of clover_exchange (clover.f90:348-500)
    SUBROUTINE clover_b200_exchange(fields, depth) BIND(C, NAME='clover_b200_exchange_')
      IMPORT :: C_INT
      INTEGER(C_INT) :: fields(15), depth
    END SUBROUTINE clover_b200_exchange
    ! replaces clover_min (clover.f90:3641-3657) / the five clover_sum calls of field_summary.f90:120-124
    SUBROUTINE clover_b200_min(value) BIND(C, NAME='clover_b200_min_')
      IMPORT :: C_DOUBLE
      REAL(C_DOUBLE) :: value
    END SUBROUTINE clover_b200_min
    SUBROUTINE clover_b200_sum(values, n) BIND(C, NAME='clover_b200_sum_')
      IMPORT :: C_INT, C_DOUBLE
      REAL(C_DOUBLE) :: values(*)
      INTEGER(C_INT) :: n
    END SUBROUTINE clover_b200_sum
    SUBROUTINE timer_c(elapsed_time) BIND(C, NAME='timer_c_')
      IMPORT :: C_DOUBLE
      REAL(C_DOUBLE) :: elapsed_time
    END SUBROUTINE timer_c
'''


def main():
    print("! clover_b200_kernels.f90 -- ISO_C_BINDING interfaces of libclover_b200.so (generated by")
    print("! include/gen_fortran_interface.py from cloverleaf_b200/abi.py; do not edit).  The C symbols are the reference's")
    print("! own `*_kernel_c_` names, so a build that keeps the implicit-interface calls of CloverLeaf_ref links unchanged;")
    print("! USE this module in the L1 wrappers to get the calls type-checked.  See INTEGRATION.md.")
    print("MODULE clover_b200_kernels")
    print("  USE, INTRINSIC :: ISO_C_BINDING")
    print("  IMPLICIT NONE")
    print("  INTERFACE")
    for sym, spec in abi.KERNELS.items():
        for line in block(sym, spec):
            print(line)
    print(EXTENSION.rstrip("\n"))
    print("  END INTERFACE")
    print("END MODULE clover_b200_kernels")


if __name__ == "__main__":
    main()
