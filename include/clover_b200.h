/* clover_b200.h -- C ABI of libclover_b200.so, the B200 (sm_100a) kernel layer for CloverLeaf.
 *
 * Part 1 is the drop-in boundary: the SAME Fortran-callable symbols that
 * CloverLeaf_ref/kernels/ *_kernel_c.c export and that the L1 wrappers of CloverLeaf_ref call on
 * the `use_c_kernels` path (implicit-interface external calls: lowercase name + trailing
 * underscore, every argument by reference, INTEGER -> int*, REAL(KIND=8) -> double*, arrays ->
 * pointer to the first element of the contiguous Fortran allocation whose lower bound is
 * (x_min-2, y_min-2)).  All return void.  Array arguments are HOST pointers; the library keeps a
 * device-resident mirror of every array it has seen, keyed by the host address (Fortran allocates
 * each field once, build_field.f90:33-94, and never moves it).
 *
 * Part 2 is the small extension a GPU backend needs and the reference ABI has no word for:
 * residency control, host synchronisation, and the replacement of clover_exchange / clover_min /
 * clover_sum (clover.f90:348-500, :3621-3709) by the library's own kernels over peer memory
 * (NVLink / NVSwitch, cudaIpc-mapped exchange blocks); NCCL bootstraps the ranks and is the fallback
 * transport (ncclSend/ncclRecv/ncclAllReduce) where cudaIpc is unavailable or CLOVER_B200_P2P=0.
 *
 * Errors: the reference ABI has no status channel.  Any CUDA / NCCL failure prints a diagnostic
 * to stderr and calls abort(); nothing here ever falls back to a CPU path.
 */
#ifndef CLOVER_B200_H
#define CLOVER_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ */
/* Part 1: the reference's kernel entry points (file:line = the reference definition replaced) */

/* kernels/ideal_gas_kernel_c.c:30   called from ideal_gas.f90:64,74 */
void ideal_gas_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *density,
                         double *energy, double *pressure, double *soundspeed);

/* kernels/viscosity_kernel_c.c:31   called from viscosity.f90:57 */
void viscosity_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *celldx, double *celldy,
                         double *density0, double *pressure, double *viscosity, double *xvel0,
                         double *yvel0);

/* kernels/calc_dt_kernel_c.c:31     called from calc_dt.f90:83.  dt_min (work_array1) is scratch
 * in the reference and is NOT written here; dtlcontrol=1, jldt=kldt=1 as in the reference. */
void calc_dt_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *g_small, double *g_big,
                       double *dtmin, double *dtc_safe, double *dtu_safe, double *dtv_safe,
                       double *dtdiv_safe, double *xarea, double *yarea, double *cellx, double *celly,
                       double *celldx, double *celldy, double *volume, double *density0,
                       double *energy0, double *pressure, double *viscosity, double *soundspeed,
                       double *xvel0, double *yvel0, double *dt_min, double *dt_min_val,
                       int *dtl_control, double *xl_pos, double *yl_pos, int *jldt, int *kldt,
                       int *small);

/* kernels/PdV_kernel_c.c:32         called from PdV.f90:89.  *prdct==0 is the predictor.
 * volume_change (work_array1) is scratch in the reference and is NOT written here. */
void pdv_kernel_c_(int *prdct, int *xmin, int *xmax, int *ymin, int *ymax, double *dt, double *xarea,
                   double *yarea, double *volume, double *density0, double *density1, double *energy0,
                   double *energy1, double *pressure, double *viscosity, double *xvel0, double *xvel1,
                   double *yvel0, double *yvel1, double *volume_change);

/* kernels/revert_kernel_c.c:32      called from revert.f90:53 */
void revert_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *density0, double *density1,
                      double *energy0, double *energy1);

/* kernels/accelerate_kernel_c.c:30  called from accelerate.f90:64 */
void accelerate_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *dt, double *xarea,
                          double *yarea, double *volume, double *density0, double *pressure,
                          double *viscosity, double *xvel0, double *yvel0, double *xvel1,
                          double *yvel1);

/* kernels/flux_calc_kernel_c.c:29   called from flux_calc.f90:62 */
void flux_calc_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *dt, double *xarea,
                         double *yarea, double *xvel0, double *yvel0, double *xvel1, double *yvel1,
                         double *vol_flux_x, double *vol_flux_y);

/* kernels/advec_cell_kernel_c.c:30  called from advec_cell_driver.f90:59.  The seven work arrays
 * are scratch in the reference; the device kernels keep those intermediates on chip and do NOT
 * write them. */
void advec_cell_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, int *dir, int *sweep_number,
                          double *vertexdx, double *vertexdy, double *volume, double *density1,
                          double *energy1, double *mass_flux_x, double *vol_flux_x,
                          double *mass_flux_y, double *vol_flux_y, double *pre_vol, double *post_vol,
                          double *pre_mass, double *post_mass, double *advec_vol, double *post_ener,
                          double *ener_flux);

/* kernels/advec_mom_kernel_c.c:32   called from advec_mom_driver.f90:85,108.  Six scratch work
 * arrays, NOT written (see advec_cell). */
void advec_mom_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *vel1,
                         double *mass_flux_x, double *vol_flux_x, double *mass_flux_y,
                         double *vol_flux_y, double *volume, double *density1, double *node_flux,
                         double *node_mass_post, double *node_mass_pre, double *mom_flux,
                         double *pre_vol, double *post_vol, double *celldx, double *celldy,
                         int *which_vel, int *sweep_number, int *direction);

/* kernels/reset_field_kernel_c.c:30 called from reset_field.f90:63 */
void reset_field_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *density0,
                           double *density1, double *energy0, double *energy1, double *xvel0,
                           double *xvel1, double *yvel0, double *yvel1);

/* kernels/update_halo_kernel_c.c:32 called from update_halo.f90:86 */
void update_halo_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, int *chunk_neighbours,
                           int *tile_neighbours, double *density0, double *energy0, double *pressure,
                           double *viscosity, double *soundspeed, double *density1, double *energy1,
                           double *xvel0, double *yvel0, double *xvel1, double *yvel1,
                           double *vol_flux_x, double *vol_flux_y, double *mass_flux_x,
                           double *mass_flux_y, int *fields, int *depth);

/* kernels/field_summary_kernel_c.c:30 called from field_summary.f90:91 */
void field_summary_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *volume,
                             double *density0, double *energy0, double *pressure, double *xvel0,
                             double *yvel0, double *vol, double *mass, double *ie, double *ke,
                             double *press);

/* kernels/pack_kernel_c.c:29,81,133,185,237,288,339,390  called from clover.f90:698-3605.
 * `buffer` is the HOST communication buffer (clover.f90:329-342); its device mirror grows on
 * demand.  These eight exist for drop-in completeness and A/B tests; the fast path is
 * clover_b200_exchange_ below, which never touches host buffers. */
void clover_pack_message_left_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *field,
                                 double *buffer, int *cell_data, int *vertex_data, int *x_face_data,
                                 int *y_face_data, int *depth, int *field_type, int *buffer_offset);
void clover_unpack_message_left_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *field,
                                   double *buffer, int *cell_data, int *vertex_data,
                                   int *x_face_data, int *y_face_data, int *depth, int *field_type,
                                   int *buffer_offset);
void clover_pack_message_right_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *field,
                                  double *buffer, int *cell_data, int *vertex_data, int *x_face_data,
                                  int *y_face_data, int *depth, int *field_type, int *buffer_offset);
void clover_unpack_message_right_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *field,
                                    double *buffer, int *cell_data, int *vertex_data,
                                    int *x_face_data, int *y_face_data, int *depth, int *field_type,
                                    int *buffer_offset);
void clover_pack_message_top_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *field,
                                double *buffer, int *cell_data, int *vertex_data, int *x_face_data,
                                int *y_face_data, int *depth, int *field_type, int *buffer_offset);
void clover_unpack_message_top_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *field,
                                  double *buffer, int *cell_data, int *vertex_data, int *x_face_data,
                                  int *y_face_data, int *depth, int *field_type, int *buffer_offset);
void clover_pack_message_bottom_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *field,
                                   double *buffer, int *cell_data, int *vertex_data,
                                   int *x_face_data, int *y_face_data, int *depth, int *field_type,
                                   int *buffer_offset);
void clover_unpack_message_bottom_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *field,
                                     double *buffer, int *cell_data, int *vertex_data,
                                     int *x_face_data, int *y_face_data, int *depth, int *field_type,
                                     int *buffer_offset);

/* kernels/initialise_chunk_kernel_c.c:29 called from initialise_chunk.f90:61 */
void initialise_chunk_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *min_x,
                                double *min_y, double *dx, double *dy, double *vertexx,
                                double *vertexdx, double *vertexy, double *vertexdy, double *cellx,
                                double *celldx, double *celly, double *celldy, double *volume,
                                double *xarea, double *yarea);

/* kernels/generate_chunk_kernel_c.c:33 called from generate_chunk.f90:77 */
void generate_chunk_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *vertexx,
                              double *vertexy, double *cellx, double *celly, double *density0,
                              double *energy0, double *xvel0, double *yvel0, int *number_of_states,
                              double *state_density, double *state_energy, double *state_xvel,
                              double *state_yvel, double *state_xmin, double *state_xmax,
                              double *state_ymin, double *state_ymax, double *state_radius,
                              int *state_geometry, int *g_rect, int *g_circ, int *g_point);

/* timer_c.c:32  called from timer.f90:31 (host wall clock; kept so the library links in place
 * of the reference's C objects) */
void timer_c_(double *elapsed_time);

/* ------------------------------------------------------------------------------------------ */
/* Part 2: GPU-backend extension (no reference counterpart unless cited) */

/* Select the CUDA device and create the streams.  Optional: the first kernel call does it on
 * device 0 (or $CLOVER_B200_DEVICE).  Aborts if no sm_100 device is present. */
void clover_b200_init_(int *device);
/* Free every device mirror, the streams and the communicator. */
void clover_b200_finalize_(void);

/* Residency.  *on != 0 (default): arrays are uploaded the first time their host address is seen
 * and then live on the device; results stay there until clover_b200_sync_to_host_.  *on == 0:
 * copy-in / copy-out -- every call uploads its inputs from the host arrays and downloads its
 * outputs before returning (a literal drop-in, used by the per-kernel A/B tests and the
 * host-buffer `e2e` measurement). */
void clover_b200_set_resident_(int *on);
/* Deferred execution (resident mode only).  Kernel entry points record their call and return; the
 * recorded stretch runs -- in order -- as soon as a result has to reach the host (calc_dt's dt,
 * field_summary's sums, any download / pack / sync below).  With fusion on (default, or
 * $CLOVER_B200_FUSE=0 to disable) recognised runs of calls execute as one kernel each:
 * ideal_gas+viscosity+calc_dt, PdV predictor+ideal_gas+revert, accelerate+PdV corrector+flux_calc,
 * advec_mom x/y velocity pairs, and reset_field / revert become buffer swaps / lazy copies.  Array
 * contents at every host-observable point are bit-identical with fusion on or off. */
void clover_b200_set_fusion_(int *on);
/* 1 (default): the fused launches stream their inputs through TMA tile staging (csrc/tma.cuh); 0: the register /
 * L2-prefetch variants of the same launches (A/B profiling, tests).  Results are bit-identical either way. */
void clover_b200_set_tma_(int *on);
/* Forget all device mirrors (the host arrays were modified behind the library's back). */
void clover_b200_invalidate_(void);
/* Forget the mirror of ONE host array (call before the host frees / re-uses that address). */
void clover_b200_forget_(double *host_array);
/* Re-upload one array from its host copy (same effect as invalidate for that array only).  An address no
 * kernel has seen yet has no mirror: nothing to do, its first use uploads it.  NOTE (deferred execution): a
 * first-use upload happens when the recorded stretch RUNS, not when the kernel entry point was called -- host
 * code that writes an array after passing it to a kernel must call clover_b200_device_synchronize_ first
 * (CloverLeaf never does this: every array is written by kernels only after start.f90). */
void clover_b200_upload_(double *host_array);
/* Copy one array back to the host (by host address; works for the 2-D fields, the 2-D geometry volume / xarea /
 * yarea and the 1-D geometry).  An address no kernel has seen yet is left alone: the host copy is the only copy. */
void clover_b200_download_(double *host_array);
/* Copy back the fields of the registered chunk selected by the 15-entry mask (ids of
 * data.f90:51-66); this is the hook visit.f90:65-77 / a debugger needs for the hydro fields.  The 1-D geometry
 * visit.f90:127,131 also reads (vertexx, vertexy) needs no hook: initialise_chunk_kernel_c_ always leaves the
 * eight 1-D geometry arrays up to date on the host as well as on the device. */
void clover_b200_sync_to_host_(int *fields);
/* Block until all device work issued so far has finished. */
void clover_b200_device_synchronize_(void);

/* Tell the library which host arrays are the 15 exchangeable fields of this process's chunk and
 * who its neighbours are (chunk numbers, 1-based, -1 = external; order left,right,bottom,top as
 * chunk%chunk_neighbours in definitions.f90:172-201).  Needed by exchange_ and sync_to_host_. */
void clover_b200_register_chunk_(int *xmin, int *xmax, int *ymin, int *ymax, int *chunk_neighbours,
                                 double *density0, double *density1, double *energy0,
                                 double *energy1, double *pressure, double *viscosity,
                                 double *soundspeed, double *xvel0, double *xvel1, double *yvel0,
                                 double *yvel1, double *vol_flux_x, double *vol_flux_y,
                                 double *mass_flux_x, double *mass_flux_y);

/* Communicator bootstrap (replaces clover_init_comms, clover.f90:70-94).  Rank 0 calls
 * get_unique_id_ (128 bytes), the host side broadcasts it by any means (MPI_BCAST in the Fortran
 * driver, torch.distributed in bench.py), every rank calls comm_init_.  rank = chunk - 1. */
void clover_b200_comm_get_unique_id_(char *id128);
void clover_b200_comm_init_(int *nranks, int *rank, char *id128);

/* clover_exchange (clover.f90:348-500).  Default transport (csrc/halo.cu): ONE kernel packs the strips of all
 * requested fields (message layout = the reference's: per-field offset = running sum of depth*(edge+5),
 * clover.f90:368-375) straight into the face neighbours' receive slots and the depth x depth corner blocks into
 * the diagonal neighbours' slots through peer memory, publishes a sequence number (st.release.sys), waits for
 * its own neighbours' (ld.acquire.sys, bounded: a peer that never arrives makes the kernel trap and the host
 * abort with a diagnostic) and unpacks.  The reference obtains the corners by running left/right before
 * bottom/top; the values that land in the corner cells are identical (tests/test_dist.py, bench.py `parity`).
 * Fallback transport: device pack -> ncclSend/ncclRecv (left/right phase, then bottom/top) -> device unpack. */
void clover_b200_exchange_(int *fields, int *depth);
/* clover_min (clover.f90:3640-3656): minimum of one double over all ranks, result on every rank.  Called right
 * after calc_dt_kernel_c_ (timestep.f90:86-90) it costs no launch: the calc_dt launch has already folded the
 * minimum over all ranks (peer-memory mailboxes, rank order) and this call returns min(*value, that minimum) --
 * which is the all-rank minimum of *value provided *value = MIN(dt_min_val, bounds that are the same on every
 * rank), as in timestep.f90; a *value above the dt_min_val calc_dt returned aborts with a message.  In any other
 * position it is a stand-alone all-reduce over the same mailboxes (fallback transport: ncclAllReduce(min)). */
void clover_b200_min_(double *value);
/* clover_sum (clover.f90:3621-3637), n values at once, result on every rank (the reference reduces to rank 0
 * only); folded in rank order, so every rank holds bit-identical sums (fallback ncclAllReduce(sum)).  Called with
 * the five sums field_summary_kernel_c_ has just returned (in the kernel's argument order: vol, mass, ie, ke,
 * press) it costs no launch either: the field_summary launch folded them over all ranks already. */
void clover_b200_sum_(double *values, int *n);

/* Accounting for bench.py: kernels launched so far by this library, and (when enabled with
 * *on != 0) per-kernel CUDA-event timing accumulated by name. */
void clover_b200_launch_count_(long long *n);
void clover_b200_profile_(int *on);
/* Writes up to *max entries; names are 32-byte NUL-padded records; returns the number in *n. */
void clover_b200_profile_get_(int *max, char *names32, double *total_ms, long long *calls, int *n);
void clover_b200_profile_reset_(void);
/* In-situ timeline: while on, every launch records %globaltimer stamps (first CTA start, last CTA end, begin and
 * end of its wait for the preceding halo kernel / for the neighbours' strips) WITHOUT serialising the launches;
 * trace_dump_ writes them as CSV (index,name,start_ns,end_ns,wait_begin_ns,wait_end_ns) and switches tracing off. */
void clover_b200_trace_(int *on);
void clover_b200_trace_dump_(const char *path);
/* Bytes copied host->device and device->host so far (all modes). */
void clover_b200_copy_bytes_(long long *h2d, long long *d2h);
/* Bytes this rank has written into its neighbours' memory (halo strips + corner blocks) and the number of
 * exchanges so far; bench.py reports the halo traffic against NVLink bandwidth from these. */
void clover_b200_halo_bytes_(long long *bytes, long long *exchanges);
/* Which transport the multi-rank data path uses: *p2p = 1 peer memory (default), 0 the NCCL fallback / one rank. */
void clover_b200_transport_(int *p2p);

/* Self-test of the library's branch-free fp64 div / rcp / sqrt against the compiler's IEEE operators
 * on *n pseudo-random + adversarial operand pairs: *mismatches must come back 0; *flagged counts
 * operands the fast sequence hands to the generic path (out of its guarded range). */
void clover_b200_selftest_math_(long long *n, long long *seed, long long *mismatches, long long *flagged,
                                long long *checked);

/* Device-side timing for bench.py: record event `slot` (0..7) on the library's stream; elapsed
 * milliseconds between two recorded slots (waits for the later one). */
void clover_b200_event_record_(int *slot);
void clover_b200_event_elapsed_ms_(int *slot_a, int *slot_b, double *ms);
/* Page-lock / unlock a host array (optional; uploads and downloads then run at full PCIe rate). */
void clover_b200_pin_(double *host_array, long long *bytes);
void clover_b200_unpin_(double *host_array);

#ifdef __cplusplus
}
#endif
#endif /* CLOVER_B200_H */
