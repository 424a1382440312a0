// advec.cu -- the advective remap: advec_cell and advec_mom, x and y sweeps, fp64 CUDA for sm_100a.
//
// The reference C kernels stage seven (cell) / six (momentum) full-size work arrays through memory
// (advec_cell_kernel_c.c, advec_mom_kernel_c.c: 22-23 array passes per call).  Here every sweep is
// ONE fused kernel that keeps pre/post volumes, node fluxes, node masses and the limited fluxes
// on chip and touches each field once: advec_cell 8 (sweep 1) / 7 (sweep 2) passes, advec_mom 7 / 6
// passes per velocity component.
//
//   x sweeps: the stencil runs along the unit-stride direction.  One warp owns a run of consecutive
//             cells / nodes of one row; neighbour values computed by the adjacent lane (pre-volume,
//             face flux, node mass, momentum flux) travel by warp shuffle, and runs overlap by one
//             (cell) or two (momentum) lanes so no lane ever recomputes an expensive flux.
//   y sweeps: one thread owns a column segment and marches along k with a rolling register window,
//             so each value is loaded once, every access is coalesced along j, and the flux through
//             face k+1 computed in one iteration is the flux through face k of the next.
//
// The updates are out of place (density1/energy1/vel1 are read at j-2..j+2 or k-2..k+2 by other
// threads): each kernel reads the current device buffer of the field and writes its second buffer
// in full (loop range = new values, everything else copied), then the runtime swaps the two.
//
// Numerics: -fmad=false, evaluation order of the reference C source => bit-identical fields.
#include "clover_b200.h"
#include "common.cuh"
#include "advec.cuh"

namespace clv {

// ================================================================================================
// advec_cell, x sweep.  Warp = 32 consecutive j of one row; lanes 0..30 own their cell and left face.
constexpr int ACX_ROWS = 8;  // warps (rows) per block

template <int SWEEP>
__device__ __forceinline__ double cell_pre_vol_x(int pitch, int j, int k, const double* __restrict__ volume,
                                                 const double* __restrict__ vfx,
                                                 const double* __restrict__ vfy) {
  const size_t c = idx2(pitch, j, k);
  if (SWEEP == 1) return volume[c] + (vfx[c + 1] - vfx[c] + vfy[c + pitch] - vfy[c]);  // :77-81
  return volume[c] + vfx[c + 1] - vfx[c];                                                // :94-96
}

template <int SWEEP>
__global__ void __launch_bounds__(32 * ACX_ROWS)
    advec_cell_x_kernel(int nx, int ny, int pitch, const double* __restrict__ vertexdx,
                        const double* __restrict__ volume, const double* __restrict__ d_old,
                        const double* __restrict__ e_old, double* __restrict__ d_new,
                        double* __restrict__ e_new, double* __restrict__ mass_flux_x,
                        const double* __restrict__ vfx, const double* __restrict__ vfy) {
  const int lane = threadIdx.x;
  const int k = -1 + (int)(blockIdx.y * ACX_ROWS + threadIdx.y);
  if (k > ny + 2) return;  // warp-uniform
  const int j = -1 + (int)blockIdx.x * 31 + lane;
  const bool in_array = (j <= nx + 2);
  const bool owned = in_array && (lane < 31);
  if (k < 1 || k > ny) {  // halo rows: carry the old values over
    if (owned) {
      const size_t c = idx2(pitch, j, k);
      d_new[c] = d_old[c];
      e_new[c] = e_old[c];
    }
    return;
  }
  // stage 1: pre-sweep volume of my cell (valid cells -1..nx+2), and of my left neighbour
  const int jc = clampi(j, -1, nx + 2);
  if (k + PF_ROWS <= ny) {
    const size_t pf = idx2(pitch, jc, k + PF_ROWS);
    prefetch_l2(volume + pf); prefetch_l2(vfx + pf); prefetch_l2(vfy + pf); prefetch_l2(d_old + pf);
    prefetch_l2(e_old + pf);
  }
  const double pv = cell_pre_vol_x<SWEEP>(pitch, jc, k, volume, vfx, vfy);
  double pv_left = __shfl_up_sync(0xffffffffu, pv, 1);
  if (lane == 0) pv_left = cell_pre_vol_x<SWEEP>(pitch, clampi(j - 1, -1, nx + 2), k, volume, vfx, vfy);
  // stage 2: fluxes through my left face (valid faces 1..nx+2)
  const int jf = clampi(j, 1, nx + 2);
  const size_t f = idx2(pitch, jf, k);
  const double vf = vfx[f];
  const double dm2 = d_old[f - 2], dm1 = d_old[f - 1], d0 = d_old[f];
  const double em2 = e_old[f - 2], em1 = e_old[f - 1], e0 = e_old[f];
  const int jup = (jf + 1 < nx + 2) ? jf + 1 : nx + 2;  // MIN(j+1,x_max+2), :114
  const double dp1 = d_old[idx2(pitch, jup, k)], ep1 = e_old[idx2(pitch, jup, k)];
  double mf, ef;
  {
    const bool pos = vf > 0.0;
    const double pvd = pos ? pv_left : pv;
    const double vdf = vertexdx[jf + 1];
    const double vdd = pos ? vertexdx[jf] : vertexdx[jup + 1];
    cell_face_flux(vf, pvd, pos ? dm2 : dp1, pos ? dm1 : d0, pos ? d0 : dm1, pos ? em2 : ep1,
                   pos ? em1 : e0, pos ? e0 : em1, vdf, vdd, mf, ef);
  }
  const double mf_right = __shfl_down_sync(0xffffffffu, mf, 1);
  const double ef_right = __shfl_down_sync(0xffffffffu, ef, 1);
  const double vf_right = __shfl_down_sync(0xffffffffu, vf, 1);
  if (!owned) return;
  const size_t c = idx2(pitch, j, k);
  if (j >= 1) mass_flux_x[c] = mf;  // faces 1..nx+2
  if (j >= 1 && j <= nx) {
    // stage 3 (:156-175)
    const double pre_mass = d0 * pv;
    const double post_mass = pre_mass + mf - mf_right;
    const double post_ener = (e0 * pre_mass + ef - ef_right) / post_mass;
    const double advec_vol = pv + vf - vf_right;
    d_new[c] = post_mass / advec_vol;
    e_new[c] = post_ener;
  } else {
    d_new[c] = d_old[c];
    e_new[c] = e_old[c];
  }
}

// ================================================================================================
// advec_cell, y sweep.  Thread = one column, marching over a segment of rows.
constexpr int ACY_THREADS = 128;
constexpr int ACY_SEG = 32;  // rows per segment

template <int SWEEP>
__device__ __forceinline__ double cell_pre_vol_y(double vol, double fy0, double fy1, double fx0, double fx1) {
  if (SWEEP == 1) return vol + (fy1 - fy0 + fx1 - fx0);  // :189-193
  return vol + fy1 - fy0;                                // :207-209
}

template <int SWEEP>
__global__ void __launch_bounds__(ACY_THREADS)
    advec_cell_y_kernel(int nx, int ny, int pitch, const double* __restrict__ vertexdy,
                        const double* __restrict__ volume, const double* __restrict__ d_old,
                        const double* __restrict__ e_old, double* __restrict__ d_new,
                        double* __restrict__ e_new, double* __restrict__ mass_flux_y,
                        const double* __restrict__ vfx, const double* __restrict__ vfy) {
  const int j = -XOFF + (int)(blockIdx.x * ACY_THREADS + threadIdx.x);
  if (j < -1 || j > nx + 2) return;
  const int ks = 1 + (int)blockIdx.y * ACY_SEG;
  const int ke = (ks + ACY_SEG - 1 < ny) ? ks + ACY_SEG - 1 : ny;
  const bool first = (blockIdx.y == 0), last = (ke == ny);
  // halo rows and halo columns: carry the old values over
  if (first)
    for (int k = -1; k <= 0; ++k) {
      const size_t c = idx2(pitch, j, k);
      d_new[c] = d_old[c];
      e_new[c] = e_old[c];
    }
  if (last)
    for (int k = ny + 1; k <= ny + 2; ++k) {
      const size_t c = idx2(pitch, j, k);
      d_new[c] = d_old[c];
      e_new[c] = e_old[c];
    }
  if (j < 1 || j > nx) {
    for (int k = ks; k <= ke; ++k) {
      const size_t c = idx2(pitch, j, k);
      d_new[c] = d_old[c];
      e_new[c] = e_old[c];
    }
    return;
  }
  // rolling window, entering iteration k: d/e at k-2..k+1, pre_vol at k-1,k, vfy at k,k+1, flux at k
  size_t c = idx2(pitch, j, ks);
  const size_t P = (size_t)pitch;
  double dm2 = d_old[c - 2 * P], dm1 = d_old[c - P], d0 = d_old[c], dp1 = d_old[c + P];
  double em2 = e_old[c - 2 * P], em1 = e_old[c - P], e0 = e_old[c], ep1 = e_old[c + P];
  double fy0 = vfy[c], fy1 = vfy[c + P];
  double pvm1 = cell_pre_vol_y<SWEEP>(volume[c - P], vfy[c - P], fy0, vfx[c - P], vfx[c - P + 1]);
  double pv0 = cell_pre_vol_y<SWEEP>(volume[c], fy0, fy1, vfx[c], vfx[c + 1]);
  double mf0, ef0;
  {
    const bool pos = fy0 > 0.0;  // face ks: upwind ks-2, donor ks-1 | upwind ks+1, donor ks
    const double vdf = vertexdy[ks + 1];
    const double vdd = pos ? vertexdy[ks] : vertexdy[ks + 2];
    cell_face_flux(fy0, pos ? pvm1 : pv0, pos ? dm2 : dp1, pos ? dm1 : d0, pos ? d0 : dm1,
                   pos ? em2 : ep1, pos ? em1 : e0, pos ? e0 : em1, vdf, vdd, mf0, ef0);
    if (first) mass_flux_y[c] = mf0;
  }
  const int kend = last ? ny + 1 : ke;  // the last segment also produces face ny+2
  for (int k = ks; k <= kend; ++k, c += P) {
    // new row k+2 (clamped at ny+2: MIN(k+1,y_max+2) of :227 for the face k+1)
    const int k2 = (k + 2 < ny + 2) ? k + 2 : ny + 2;
    const size_t c2 = idx2(pitch, j, k2);
    if (k + PF_MARCH <= ny + 2) {
      const size_t pf = c + (size_t)PF_MARCH * P;
      prefetch_l2(d_old + pf); prefetch_l2(e_old + pf); prefetch_l2(vfy + pf); prefetch_l2(volume + pf);
      if (SWEEP == 1) prefetch_l2(vfx + pf);
    }
    const double dp2 = d_old[c2], ep2 = e_old[c2];
    const double fy2 = vfy[c + 2 * P];
    const double pv1 = cell_pre_vol_y<SWEEP>(volume[c + P], fy1, fy2, vfx[c + P], vfx[c + P + 1]);
    // face k+1: flux>0: upwind k-1, donor k, downwind k+1 ; else upwind min(k+2,ny+2), donor k+1, downwind k
    double mf1, ef1;
    {
      const bool pos = fy1 > 0.0;
      const double vdf = vertexdy[k + 2];
      const double vdd = pos ? vertexdy[k + 1] : vertexdy[k2 + 1];
      cell_face_flux(fy1, pos ? pv0 : pv1, pos ? dm1 : dp2, pos ? d0 : dp1, pos ? dp1 : d0,
                     pos ? em1 : ep2, pos ? e0 : ep1, pos ? ep1 : e0, vdf, vdd, mf1, ef1);
    }
    mass_flux_y[c + P] = mf1;
    if (k <= ny) {
      // :266-286
      const double pre_mass = d0 * pv0;
      const double post_mass = pre_mass + mf0 - mf1;
      const double post_ener = (e0 * pre_mass + ef0 - ef1) / post_mass;
      const double advec_vol = pv0 + fy0 - fy1;
      d_new[c] = post_mass / advec_vol;
      e_new[c] = post_ener;
    }
    dm2 = dm1; dm1 = d0; d0 = dp1; dp1 = dp2;
    em2 = em1; em1 = e0; e0 = ep1; ep1 = ep2;
    fy0 = fy1; fy1 = fy2;
    pvm1 = pv0; pv0 = pv1;
    mf0 = mf1; ef0 = ef1;
  }
}

// ================================================================================================
// advec_mom.  post_vol of a cell by mom_sweep (advec_mom_kernel_c.c:69-121); pre_vol is not needed:
// node_mass_pre is derived from node_mass_post and the node fluxes (:151-158).
template <int MOMSWEEP>
__device__ __forceinline__ double cell_post_mass(int pitch, size_t c, const double* __restrict__ density1,
                                                 const double* __restrict__ volume,
                                                 const double* __restrict__ vfx,
                                                 const double* __restrict__ vfy) {
  double post_vol;
  if (MOMSWEEP == 1) post_vol = volume[c] + vfy[c + pitch] - vfy[c];
  else if (MOMSWEEP == 2) post_vol = volume[c] + vfx[c + 1] - vfx[c];
  else post_vol = volume[c];
  return density1[c] * post_vol;
}

// ---- x sweep: warp = 32 consecutive nodes of one row, lanes 1..30 own their node -----------------
constexpr int AMX_ROWS = 8;

template <int MOMSWEEP, int NVEL>
__global__ void __launch_bounds__(32 * AMX_ROWS)
    advec_mom_x_kernel(int nx, int ny, int pitch, const double* __restrict__ celldx,
                       const double* __restrict__ volume, const double* __restrict__ density1,
                       const double* __restrict__ mfx, const double* __restrict__ vfx,
                       const double* __restrict__ vfy, const double* __restrict__ v0_old,
                       double* __restrict__ v0_new, const double* __restrict__ v1_old,
                       double* __restrict__ v1_new) {
  const int lane = threadIdx.x;
  const int k = -1 + (int)(blockIdx.y * AMX_ROWS + threadIdx.y);
  if (k > ny + 3) return;
  const int j = -2 + (int)blockIdx.x * 30 + lane;
  const bool owned = (lane >= 1) && (lane <= 30) && (j >= -1) && (j <= nx + 3);
  const double* vold[2] = {v0_old, v1_old};
  double* vnew[2] = {v0_new, v1_new};
  if (k < 1 || k > ny + 1) {
    if (owned) {
      const size_t n = idx2(pitch, j, k);
#pragma unroll
      for (int v = 0; v < NVEL; ++v) vnew[v][n] = vold[v][n];
    }
    return;
  }
  // node quantities: node_flux is needed for j in -1..nx+2, node masses for 0..nx+2; lanes outside
  // are clamped onto valid memory and their results are never committed
  const int jn = clampi(j, 0, nx + 2);
  const size_t n = idx2(pitch, jn, k);
  const size_t P = (size_t)pitch;
  if (k + PF_ROWS <= ny + 1) {
    const size_t pf = n + (size_t)PF_ROWS * P;
    prefetch_l2(volume + pf); prefetch_l2(density1 + pf); prefetch_l2(mfx + pf); prefetch_l2(v0_old + pf);
    if (MOMSWEEP == 1) prefetch_l2(vfy + pf);
    if (NVEL == 2) prefetch_l2(v1_old + pf);
  }
  // node_flux (:124-134) and its left neighbour
  const size_t nfi = idx2(pitch, clampi(j, -1, nx + 2), k);
  const double nf = 0.25 * (mfx[nfi - P] + mfx[nfi] + mfx[nfi - P + 1] + mfx[nfi + 1]);
  double nf_left = __shfl_up_sync(0xffffffffu, nf, 1);
  if (lane == 0) {
    const size_t m = idx2(pitch, clampi(j - 1, -1, nx + 2), k);
    nf_left = 0.25 * (mfx[m - P] + mfx[m] + mfx[m - P + 1] + mfx[m + 1]);
  }
  // node_mass_post (:135-150): cells (j,k-1),(j,k),(j-1,k-1),(j-1,k) in that order
  const double nm_post = 0.25 * (cell_post_mass<MOMSWEEP>(pitch, n - P, density1, volume, vfx, vfy) +
                                 cell_post_mass<MOMSWEEP>(pitch, n, density1, volume, vfx, vfy) +
                                 cell_post_mass<MOMSWEEP>(pitch, n - P - 1, density1, volume, vfx, vfy) +
                                 cell_post_mass<MOMSWEEP>(pitch, n - 1, density1, volume, vfx, vfy));
  const double nm_pre = nm_post - nf_left + nf;  // :151-158
  const double nm_pre_right = __shfl_down_sync(0xffffffffu, nm_pre, 1);
  // mom_flux through node jn (:159-189), valid for lanes 0..30
  const bool neg = nf < 0.0;
  const int jup = clampi(neg ? jn + 2 : jn - 1, -1, nx + 3);
  const int jdon = neg ? jn + 1 : jn;
  const int jdown = neg ? jn : jn + 1;
  const int jdif = clampi(neg ? jdon : jup, -1, nx + 2);
  const double width = celldx[jn + 1], width_dif = celldx[jdif + 1];
  const double nmp_don = neg ? nm_pre_right : nm_pre;
#pragma unroll
  for (int v = 0; v < NVEL; ++v) {
    const double* __restrict__ vel = vold[v];
    const double v_up = vel[idx2(pitch, jup, k)];
    const double v_don = vel[idx2(pitch, jdon, k)];
    const double v_down = vel[idx2(pitch, jdown, k)];
    const double mom_flux = mom_face_flux(nf, nmp_don, v_up, v_don, v_down, width, width_dif);
    const double mom_flux_left = __shfl_up_sync(0xffffffffu, mom_flux, 1);
    if (owned) {
      const size_t o = idx2(pitch, j, k);
      if (j >= 1 && j <= nx + 1)
        vnew[v][o] = ddiv(vel[o] * nm_pre + mom_flux_left - mom_flux, nm_post);  // :191-201
      else
        vnew[v][o] = vel[o];
    }
  }
}

// ---- y sweep: thread = one node column, marching over a segment of rows --------------------------
constexpr int AMY_THREADS = 128;
constexpr int AMY_SEG = 32;

template <int MOMSWEEP, int NVEL>
__global__ void __launch_bounds__(AMY_THREADS)
    advec_mom_y_kernel(int nx, int ny, int pitch, const double* __restrict__ celldy,
                       const double* __restrict__ volume, const double* __restrict__ density1,
                       const double* __restrict__ mfy, const double* __restrict__ vfx,
                       const double* __restrict__ vfy, const double* __restrict__ v0_old,
                       double* __restrict__ v0_new, const double* __restrict__ v1_old,
                       double* __restrict__ v1_new) {
  const int j = -XOFF + (int)(blockIdx.x * AMY_THREADS + threadIdx.x);
  if (j < -1 || j > nx + 3) return;
  const int ks = 1 + (int)blockIdx.y * AMY_SEG;
  const int ke = (ks + AMY_SEG - 1 < ny + 1) ? ks + AMY_SEG - 1 : ny + 1;
  const bool first = (blockIdx.y == 0), last = (ke == ny + 1);
  const double* vold[2] = {v0_old, v1_old};
  double* vnew[2] = {v0_new, v1_new};
  const size_t P = (size_t)pitch;
  if (first)
    for (int k = -1; k <= 0; ++k)
#pragma unroll
      for (int v = 0; v < NVEL; ++v) vnew[v][idx2(pitch, j, k)] = vold[v][idx2(pitch, j, k)];
  if (last)
    for (int k = ny + 2; k <= ny + 3; ++k)
#pragma unroll
      for (int v = 0; v < NVEL; ++v) vnew[v][idx2(pitch, j, k)] = vold[v][idx2(pitch, j, k)];
  if (j < 1 || j > nx + 1) {
    for (int k = ks; k <= ke; ++k)
#pragma unroll
      for (int v = 0; v < NVEL; ++v) vnew[v][idx2(pitch, j, k)] = vold[v][idx2(pitch, j, k)];
    return;
  }
  // ---- set-up for row k = ks-1 -------------------------------------------------------------------
  size_t c = idx2(pitch, j, ks - 1);
  // mass_flux_y rows ks-2, ks-1, ks at columns j-1, j
  const double ma0 = mfy[c - P - 1], mb0 = mfy[c - P];  // row ks-2
  double ma1 = mfy[c - 1], mb1 = mfy[c];          // row ks-1
  double ma2 = mfy[c + P - 1], mb2 = mfy[c + P];  // row ks
  const double nf_m2 = 0.25 * (ma0 + mb0 + ma1 + mb1);  // node_flux(ks-2)  (:205-214)
  double nf0 = 0.25 * (ma1 + mb1 + ma2 + mb2);          // node_flux(ks-1)
  // cell post-masses rows ks-2, ks-1 at columns j-1 (L) and j (R)
  double cmL = cell_post_mass<MOMSWEEP>(pitch, c - 1, density1, volume, vfx, vfy);      // (j-1, ks-1)
  double cmR = cell_post_mass<MOMSWEEP>(pitch, c, density1, volume, vfx, vfy);          // (j  , ks-1)
  double nmpost0;
  {
    const double cmLm = cell_post_mass<MOMSWEEP>(pitch, c - P - 1, density1, volume, vfx, vfy);
    const double cmRm = cell_post_mass<MOMSWEEP>(pitch, c - P, density1, volume, vfx, vfy);
    // node_mass_post(ks-1): (j,k-1),(j,k),(j-1,k-1),(j-1,k)  (:216-231)
    nmpost0 = 0.25 * (cmRm + cmR + cmLm + cmL);
  }
  double nmpre0 = nmpost0 - nf_m2 + nf0;  // :232-239
  ma1 = ma2; mb1 = mb2;                   // now (ma1,mb1) = mass_flux_y row k+1 for k = ks-1
  double vm1[NVEL], v0[NVEL], vp1[NVEL], mom_prev[NVEL];
#pragma unroll
  for (int v = 0; v < NVEL; ++v) {
    vm1[v] = vold[v][c - P];
    v0[v] = vold[v][c];
    vp1[v] = vold[v][c + P];
    mom_prev[v] = 0.0;
  }
  // ---- march: k = ks-1 is the priming iteration (computes mom_flux(ks-1), stores nothing) --------
  for (int k = ks - 1; k <= ke; ++k, c += P) {
    if (k + PF_MARCH <= ny + 2) {
      const size_t pf = c + (size_t)PF_MARCH * P;
      prefetch_l2(mfy + pf); prefetch_l2(density1 + pf); prefetch_l2(volume + pf); prefetch_l2(v0_old + pf);
      if (MOMSWEEP == 2) prefetch_l2(vfx + pf);
      if (NVEL == 2) prefetch_l2(v1_old + pf);
    }
    // node_flux(k+1) from mass_flux_y rows k+1 (held) and k+2 (new)
    const double na = mfy[c + 2 * P - 1], nb = mfy[c + 2 * P];
    const double nf1 = 0.25 * (ma1 + mb1 + na + nb);
    // node_mass_post(k+1) from cell rows k (held) and k+1 (new)
    const double cmL1 = cell_post_mass<MOMSWEEP>(pitch, c + P - 1, density1, volume, vfx, vfy);
    const double cmR1 = cell_post_mass<MOMSWEEP>(pitch, c + P, density1, volume, vfx, vfy);
    const double nmpost1 = 0.25 * (cmR + cmR1 + cmL + cmL1);
    const double nmpre1 = nmpost1 - nf0 + nf1;
    // mom_flux(k) (:240-270): nf<0: upwind k+2, donor k+1, downwind k, dif=donor; else upwind k-1, donor k, downwind k+1, dif=upwind
    const bool neg = nf0 < 0.0;
    const double width = celldy[k + 1];
    const double width_dif = neg ? celldy[k + 2] : celldy[k];
    const double nmp_don = neg ? nmpre1 : nmpre0;
#pragma unroll
    for (int v = 0; v < NVEL; ++v) {
      const double vp2 = vold[v][c + 2 * P];
      const double mom = mom_face_flux(nf0, nmp_don, neg ? vp2 : vm1[v], neg ? vp1[v] : v0[v],
                                       neg ? v0[v] : vp1[v], width, width_dif);
      if (k >= ks) vnew[v][c] = ddiv(v0[v] * nmpre0 + mom_prev[v] - mom, nmpost0);  // :272-282
      mom_prev[v] = mom;
      vm1[v] = v0[v]; v0[v] = vp1[v]; vp1[v] = vp2;
    }
    ma1 = na; mb1 = nb;
    nf0 = nf1;
    cmL = cmL1; cmR = cmR1;
    nmpost0 = nmpost1;
    nmpre0 = nmpre1;
  }
}

}  // namespace clv

namespace clv {

bool tma_enabled();
void run_advec_cell_tma(const Grid& g, int dir, int sweep, double* vertexdx, double* vertexdy, double* volume,
                        double* density1, double* energy1, double* mass_flux_x, double* vol_flux_x, double* mass_flux_y,
                        double* vol_flux_y);

void run_advec_cell(const Grid& g, int dir, int sweep, double* vertexdx, double* vertexdy, double* volume,
                    double* density1, double* energy1, double* mass_flux_x, double* vol_flux_x,
                    double* mass_flux_y, double* vol_flux_y) {
  // resident mode: the TMA tile kernel (advec_tma.cu); copy-in/out mode keeps the register/shuffle kernels below
  if (tma_enabled() && is_resident() && fusion_enabled()) {
    run_advec_cell_tma(g, dir, sweep, vertexdx, vertexdy, volume, density1, energy1, mass_flux_x, vol_flux_x,
                       mass_flux_y, vol_flux_y);
    return;
  }
  const double* vol = dev(g, volume, CELL, IN);
  const double* fx = dev(g, vol_flux_x, XFACE, IN);
  const double* fy = dev(g, vol_flux_y, YFACE, IN);
  const double* d_old = dev(g, density1, CELL, INOUT);
  const double* e_old = dev(g, energy1, CELL, INOUT);
  double* d_new = dev_alt(g, density1, CELL);
  double* e_new = dev_alt(g, energy1, CELL);
  if (dir == 1) {
    const double* vdx = dev(g, vertexdx, X1D_VERT, IN);
    double* mf = dev(g, mass_flux_x, XFACE, OUT);
    const dim3 grid((unsigned)((g.nx + 4 + 30) / 31), (unsigned)((g.ny + 4 + ACX_ROWS - 1) / ACX_ROWS));
    LaunchScope ls("advec_cell_x");
    if (sweep == 1)
      advec_cell_x_kernel<1><<<grid, dim3(32, ACX_ROWS), 0, stream()>>>(g.nx, g.ny, g.pitch, vdx, vol, d_old, e_old,
                                                                      d_new, e_new, mf, fx, fy);
    else
      advec_cell_x_kernel<2><<<grid, dim3(32, ACX_ROWS), 0, stream()>>>(g.nx, g.ny, g.pitch, vdx, vol, d_old, e_old,
                                                                      d_new, e_new, mf, fx, fy);
  } else {
    const double* vdy = dev(g, vertexdy, Y1D_VERT, IN);
    double* mf = dev(g, mass_flux_y, YFACE, OUT);
    const dim3 grid((unsigned)((g.nx + 3 + XOFF + ACY_THREADS) / ACY_THREADS),
                    (unsigned)((g.ny + ACY_SEG - 1) / ACY_SEG));
    LaunchScope ls("advec_cell_y");
    if (sweep == 1)
      advec_cell_y_kernel<1><<<grid, ACY_THREADS, 0, stream()>>>(g.nx, g.ny, g.pitch, vdy, vol, d_old, e_old, d_new,
                                                                e_new, mf, fx, fy);
    else
      advec_cell_y_kernel<2><<<grid, ACY_THREADS, 0, stream()>>>(g.nx, g.ny, g.pitch, vdy, vol, d_old, e_old, d_new,
                                                                e_new, mf, fx, fy);
  }
  swap_alt(density1);
  swap_alt(energy1);
}

// One launch advects vel_a and, when vel_b != nullptr, vel_b as well (advec_mom is called twice per sweep with
// identical arguments except the velocity component, advec_mom_driver.f90:85,108; node fluxes, node masses
// and the upwind choice are the same for both).
void run_advec_mom(const Grid& g, int dirn, int sweep, double* vel_a, double* vel_b, double* mass_flux_x,
                   double* vol_flux_x, double* mass_flux_y, double* vol_flux_y, double* volume, double* density1,
                   double* celldx, double* celldy) {
  const int mom_sweep = dirn + 2 * (sweep - 1);
  const double* vol = dev(g, volume, CELL, IN);
  const double* d1 = dev(g, density1, CELL, IN);
  const double* fx = dev(g, vol_flux_x, XFACE, IN);
  const double* fy = dev(g, vol_flux_y, YFACE, IN);
  const double* va_old = dev(g, vel_a, VERTEX, INOUT);
  double* va_new = dev_alt(g, vel_a, VERTEX);
  const double* vb_old = vel_b ? dev(g, vel_b, VERTEX, INOUT) : nullptr;
  double* vb_new = vel_b ? dev_alt(g, vel_b, VERTEX) : nullptr;
  if (dirn == 1) {
    const double* cdx = dev(g, celldx, X1D_CELL, IN);
    const double* mf = dev(g, mass_flux_x, XFACE, IN);
    const dim3 grid((unsigned)((g.nx + 5 + 29) / 30), (unsigned)((g.ny + 5 + AMX_ROWS - 1) / AMX_ROWS));
    const dim3 block(32, AMX_ROWS);
    LaunchScope ls(vel_b ? "advec_mom_x2" : "advec_mom_x");
#define CLV_MOMX(MS, NV) \
  advec_mom_x_kernel<MS, NV><<<grid, block, 0, stream()>>>(g.nx, g.ny, g.pitch, cdx, vol, d1, mf, fx, fy, va_old, \
                                                           va_new, vb_old, vb_new)
    if (mom_sweep == 1) { if (vel_b) CLV_MOMX(1, 2); else CLV_MOMX(1, 1); }
    else                { if (vel_b) CLV_MOMX(3, 2); else CLV_MOMX(3, 1); }
#undef CLV_MOMX
  } else {
    const double* cdy = dev(g, celldy, Y1D_CELL, IN);
    const double* mf = dev(g, mass_flux_y, YFACE, IN);
    const dim3 grid((unsigned)((g.nx + 4 + XOFF + AMY_THREADS) / AMY_THREADS),
                    (unsigned)((g.ny + 1 + AMY_SEG - 1) / AMY_SEG));
    LaunchScope ls(vel_b ? "advec_mom_y2" : "advec_mom_y");
#define CLV_MOMY(MS, NV) \
  advec_mom_y_kernel<MS, NV><<<grid, AMY_THREADS, 0, stream()>>>(g.nx, g.ny, g.pitch, cdy, vol, d1, mf, fx, fy, \
                                                                 va_old, va_new, vb_old, vb_new)
    if (mom_sweep == 2) { if (vel_b) CLV_MOMY(2, 2); else CLV_MOMY(2, 1); }
    else                { if (vel_b) CLV_MOMY(4, 2); else CLV_MOMY(4, 1); }
#undef CLV_MOMY
  }
  swap_alt(vel_a);
  if (vel_b) swap_alt(vel_b);
}

}  // namespace clv

using namespace clv;

extern "C" {

// Op::a = {vertexdx, vertexdy, volume, density1, energy1, mass_flux_x, vol_flux_x, mass_flux_y, vol_flux_y};
// iv = {dir, sweep}
void advec_cell_kernel_c_(int* xmin, int* xmax, int* ymin, int* ymax, int* dir, int* sweep_number,
                          double* vertexdx, double* vertexdy, double* volume, double* density1,
                          double* energy1, double* mass_flux_x, double* vol_flux_x,
                          double* mass_flux_y, double* vol_flux_y, double* pre_vol, double* post_vol,
                          double* pre_mass, double* post_mass, double* advec_vol, double* post_ener,
                          double* ener_flux) {
  (void)pre_vol; (void)post_vol; (void)pre_mass; (void)post_mass; (void)advec_vol; (void)post_ener; (void)ener_flux;
  Op op;
  op.kind = OP_ADVEC_CELL;
  const Grid g = op.g = grid_of_noflush(xmin, xmax, ymin, ymax);
  const int d = op.iv[0] = *dir, sweep = op.iv[1] = *sweep_number;
  if ((d != 1 && d != 2) || (sweep != 1 && sweep != 2)) fatal("advec_cell: dir=%d sweep=%d", d, sweep);
  double* a[] = {vertexdx, vertexdy, volume, density1, energy1, mass_flux_x, vol_flux_x, mass_flux_y, vol_flux_y};
  for (double* p : a) op.a[op.na++] = p;
  op.reads({vertexdx, vertexdy, volume, density1, energy1, vol_flux_x, vol_flux_y});
  op.writes({density1, energy1, d == 1 ? mass_flux_x : mass_flux_y});
  op.run = [=] {
    run_advec_cell(g, d, sweep, vertexdx, vertexdy, volume, density1, energy1, mass_flux_x, vol_flux_x, mass_flux_y,
                   vol_flux_y);
  };
  submit(std::move(op));
}

// Op::a = {vel1, mass_flux_x, vol_flux_x, mass_flux_y, vol_flux_y, volume, density1, celldx, celldy};
// iv = {which_vel, sweep, direction}
void advec_mom_kernel_c_(int* xmin, int* xmax, int* ymin, int* ymax, double* vel1, double* mass_flux_x,
                         double* vol_flux_x, double* mass_flux_y, double* vol_flux_y, double* volume,
                         double* density1, double* node_flux, double* node_mass_post,
                         double* node_mass_pre, double* mom_flux, double* pre_vol, double* post_vol,
                         double* celldx, double* celldy, int* which_vel, int* sweep_number,
                         int* direction) {
  (void)node_flux; (void)node_mass_post; (void)node_mass_pre; (void)mom_flux; (void)pre_vol; (void)post_vol;
  Op op;
  op.kind = OP_ADVEC_MOM;
  const Grid g = op.g = grid_of_noflush(xmin, xmax, ymin, ymax);
  const int dirn = *direction, sweep = *sweep_number;
  op.iv[0] = *which_vel; op.iv[1] = sweep; op.iv[2] = dirn;
  const int mom_sweep = dirn + 2 * (sweep - 1);
  if ((dirn != 1 && dirn != 2) || mom_sweep < 1 || mom_sweep > 4) fatal("advec_mom: direction=%d sweep=%d", dirn, sweep);
  double* a[] = {vel1, mass_flux_x, vol_flux_x, mass_flux_y, vol_flux_y, volume, density1, celldx, celldy};
  for (double* p : a) op.a[op.na++] = p;
  op.reads({vel1, mass_flux_x, vol_flux_x, mass_flux_y, vol_flux_y, volume, density1, celldx, celldy});
  op.writes({vel1});
  op.run = [=] {
    run_advec_mom(g, dirn, sweep, vel1, nullptr, mass_flux_x, vol_flux_x, mass_flux_y, vol_flux_y, volume, density1,
                  celldx, celldy);
  };
  submit(std::move(op));
}

}  // extern "C"
