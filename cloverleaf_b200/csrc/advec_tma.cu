// advec_tma.cu -- the advective remap with TMA tile staging (tma.cuh): persistent CTAs, the stencil neighbourhood of
// a tile arrives in shared memory as one box per field, the per-node / per-face intermediates (node fluxes, node
// masses, limited fluxes) are exchanged between the threads of a CTA through shared memory, and nothing but the
// final fields goes back to HBM.  Same arithmetic, statement for statement, as advec.cu (which remains the
// single-call path: copy-in/out mode, fusion off).
//
// advec_mom, both velocity components in one launch (advec_mom_kernel_c.c:69-286).  `s` is the sweep axis:
//   A  node_flux(s)      s = s0-2 .. s0+NS      (:124-134 / :205-214)   NS = nodes of the tile along the sweep
//      node_mass_post(s) s = s0-1 .. s0+NS      (:135-150 / :216-231)
//   B  mom_flux(s)       s = s0-1 .. s0+NS-1    (:159-189 / :240-270)   node_mass_pre = post - flux(s-1) + flux(s)
//   C  vel1(s)           s = s0   .. s0+NS-1    (:191-201 / :272-282)
// Thread grid TX x TY, RPT rows per thread.  x sweeps: the tile is (TX-4) x (TY*RPT) nodes, so that the TX threads
// of a row cover the TX-1 node fluxes, TX-3 momentum fluxes and TX-4 nodes of that row in one round each
// (plane positions 0..TX-2, 1..TX-3 and 2..TX-3).
// y sweeps: the tile is TX x (TY*RPT-4) nodes for the same reason along k.
#include "clover_b200.h"
#include "advec.cuh"
#include "common.cuh"
#include "tma.cuh"

namespace clv {

bool tma_enabled();
int sm_count();

// the halo ring (everything outside the update range 1..nx+e x 1..ny+e of the Fortran extent) is carried over to
// the new buffer by all CTAs together
__device__ __forceinline__ void ring_copy(const double* __restrict__ src, double* __restrict__ dst, int nx, int ny,
                                          int pitch, int e, int t0, int nt) {
  const int W = nx + 4 + e, H = ny + 4 + e;
  const int n_bt = 4 * W, n_lr = 4 * (H - 4);
  for (int t = t0; t < n_bt + n_lr; t += nt) {
    int j, k;
    if (t < n_bt) {
      const int r = t / W;
      j = -1 + t % W;
      k = r < 2 ? -1 + r : ny + e + (r - 1);
    } else {
      const int u = t - n_bt, c = u / (H - 4);
      k = 1 + u % (H - 4);
      j = c < 2 ? -1 + c : nx + e + (c - 1);
    }
    const size_t i = idx2(pitch, j, k);
    dst[i] = src[i];
  }
}

// ---- tile shapes of the x sweeps as build macros (the defaults are the measured optimum, profiles/
// r02_experiment_occupancy.txt, r02_experiment_timestep_rows.txt): thread rows, rows per thread, CTAs per SM.
// advec_cell x, CTAs per SM: with the flux planes inside dead boxes four fit, but 64 registers spill (94 bytes) and
// the launch is slower: 0.201 vs 0.178 ms (profiles/r02_experiment_occupancy.txt)
#ifndef CELLX_CPS
#define CELLX_CPS 3
#endif
#ifndef CELLX_TY
#define CELLX_TY 4
#endif
#ifndef CELLX_RPT
#define CELLX_RPT 2
#endif
#ifndef MOMX_RPT
#define MOMX_RPT 2
#endif
#ifndef MOMX_TY
#define MOMX_TY 4
#endif
#ifndef MOMX_CPS
#define MOMX_CPS 3
#endif

enum { MA_VOLUME = 0, MA_DENSITY1, MA_MASS_FLUX, MA_VEL_A, MA_VEL_B, MA_VOL_FLUX, MA_NARR };

// boxes a sweep stages: the post-volume of mom_sweep 1 / 2 needs a volume flux (:69-121), that of 3 / 4 does not --
// the sixth box is then neither loaded nor given room (one of eight passes less)
constexpr int mom_narr(int ms) { return ms <= 2 ? MA_NARR : MA_NARR - 1; }

template <int DIR, int TX, int TY, int RPT, int STAGES, int CPS, int MS = 1>
struct MomCfg {
  static constexpr int NT = TX * TY;
  static constexpr int ROWS = TY * RPT;
  static constexpr int W = DIR == 1 ? TX - 4 : TX;          // nodes per tile along x
  static constexpr int H = DIR == 1 ? ROWS : ROWS - 4;      // nodes per tile along y
  static constexpr int BW = DIR == 1 ? TX : TX + 4;         // box: x from j0-2
  static constexpr int BH = DIR == 1 ? H + 2 : H + 4;       // box: y from k0-1 (x sweep) / k0-2 (y sweep)
  static constexpr int OX = 2, OY = DIR == 1 ? 1 : 2;
  using Ring = TileRing<mom_narr(MS), BW, BH, STAGES>;
  static constexpr int NI = TX * ROWS;                      // one intermediate plane: thread-grid shaped
  // the two momentum-flux planes live in the volume / density1 boxes of the stage, dead after phase A
  static_assert(NI <= BW * BH, "a plane must fit a box");
  static constexpr int SMEM = Ring::BYTES + 2 * NI * 8 + 128;
  static_assert(fits_sm(SMEM, CPS), "advec_mom: CPS CTAs of this shape do not fit one SM");
};
struct MomMaps {
  CUtensorMap m[MA_NARR];
};

// MS = mom_sweep = direction + 2*(sweep-1)
template <int DIR, int MS, int TX, int TY, int RPT, int STAGES, int CPS>
__global__ void __launch_bounds__(TX* TY, CPS)
    advec_mom_tma_kernel(const __grid_constant__ MomMaps M, const double* __restrict__ va_old, double* __restrict__ va_new,
                         const double* __restrict__ vb_old, double* __restrict__ vb_new,
                         const double* __restrict__ celld, int nx, int ny, int pitch, int ntx, int ntiles,
                        const int2* __restrict__ order, Tickets tickets, int dep_start, unsigned long long* trace) {
  using Cfg = MomCfg<DIR, TX, TY, RPT, STAGES, CPS, MS>;
  constexpr int NT = Cfg::NT, W = Cfg::W, H = Cfg::H, BW = Cfg::BW, NI = Cfg::NI, OX = Cfg::OX, OY = Cfg::OY;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = align128(smem_raw);
  typename Cfg::Ring ring;
  ring.init(smem);
  double* __restrict__ s_nf = reinterpret_cast<double*>(smem + Cfg::Ring::BYTES);  // node_flux
  double* __restrict__ s_np = s_nf + NI;                                           // node_mass_post
  const int tid = threadIdx.x, lx = tid % TX, ty = tid / TX;
  const int G = gridDim.x;
  pdl_trigger();
  PdlGate gate(dep_start, trace);
  // The halo ring of the old buffers (what the preceding halo exchange / reflective boundary delivered) moves to the
  // new ones: up front, next to the first TMA loads, when this launch waits for its predecessor anyway; after the
  // tiles when the interior tiles run ahead of a halo kernel (dep_start > 0).
  if (dep_start == 0) {
    gate.need(0);
    ring_copy(va_old, va_new, nx, ny, pitch, 1, (int)blockIdx.x * NT + tid, G * NT);
    ring_copy(vb_old, vb_new, nx, ny, pitch, 1, (int)blockIdx.x * NT + tid, G * NT);
  }
  auto issue_tile = [&](int stage, int2 xy) {
    const int j0 = 1 + xy.x * W, k0 = 1 + xy.y * H;
    ring.issue(M.m, stage, j0 - OX + XOFF, k0 - OY + 1);
  };
  __shared__ int s_tile[STAGES];
  __shared__ int2 s_xy[STAGES];
  __shared__ int s_q[8];
  TileQueue<STAGES> queue(tickets, ntiles, order, s_tile, s_xy, s_q);
  const bool sched = (tid == 32);  // lane 0 of warp 1 drives the tile queue (tma.cuh)
  if (sched) queue.prime_all();
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s_tile[s] < ntiles) {
        gate.need(s_tile[s]);
        issue_tile(s, s_xy[s]);
      }
    }
  }
  // the sweep axis in box / plane coordinates: moving one node along the sweep
  constexpr int SB = DIR == 1 ? 1 : BW;   // box stride along the sweep
  constexpr int SP = DIR == 1 ? 1 : TX;   // plane stride along the sweep
  for (int it = 0;; ++it) {
    const int stage = it % STAGES;
    const int t = s_tile[stage];
    if (t >= ntiles) break;
    const int2 cur = s_xy[stage];
    gate.need(t);
    if (tid == 0) {
      const int ns = (stage + STAGES - 1) % STAGES;
      const int tn = s_tile[ns];
      if (tn < ntiles) {
        gate.need(tn);
        issue_tile(ns, s_xy[ns]);
      }
    }
    const int j0 = 1 + cur.x * W, k0 = 1 + cur.y * H;
    // celldx / celldy at the sweep positions s-1, s, s+1 of my nodes (1-D, lower bound -1 -> index s+1; clamped for
    // the nodes beyond the chunk, whose results are never stored); issued before the wait so that they overlap it
    double cw[RPT], cwm[RPT], cwp[RPT];
    {
      const int smax = (DIR == 1 ? nx : ny) + 2;
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int s_node = DIR == 1 ? j0 - 2 + lx : k0 - 2 + ty * RPT + r;
        cw[r] = celld[clampi(s_node, -1, smax) + 1];
        cwm[r] = celld[clampi(s_node - 1, -1, smax) + 1];
        cwp[r] = celld[clampi(s_node + 1, -1, smax) + 1];
      }
    }
    ring.wait(stage, (uint32_t)((it / STAGES) & 1));
    const double* svol = ring.tile(stage, MA_VOLUME);
    const double* sd1 = ring.tile(stage, MA_DENSITY1);
    double* s_ma = ring.scratch(stage, MA_VOLUME);   // mom_flux, component a: written after the barrier that ends phase A,
    double* s_mb = ring.scratch(stage, MA_DENSITY1);  // component b            the last reader of these two boxes
    const double* __restrict__ smf = ring.tile(stage, MA_MASS_FLUX);
    const double* __restrict__ sva = ring.tile(stage, MA_VEL_A);
    const double* __restrict__ svb = ring.tile(stage, MA_VEL_B);
    const double* __restrict__ svf = ring.tile(stage, MS <= 2 ? MA_VOL_FLUX : MA_VOLUME);  // (read for MS 1 / 2 only)
    // Thread (lx, ty) owns the RPT adjacent plane rows ty*RPT + r, plane position p = row*TX + lx.
    // x sweep: plane column lx <-> node j0-2+lx, plane row <-> node k0+row.
    // y sweep: plane column lx <-> node j0+lx,   plane row <-> node k0-2+row.
    // Box position of node (j,k): b = (k-k0+OY)*BW + (j-j0+OX).
    const int row0 = ty * RPT;
    const int b0 = DIR == 1 ? (row0 + OY) * BW + lx : row0 * BW + lx + OX;
    const int p0 = row0 * TX + lx;
    constexpr int NPS = DIR == 1 ? TX : Cfg::ROWS;  // plane positions along the sweep
    // ---- A: node_flux (valid at sweep positions 0..NPS-2) and node_mass_post (1..NPS-1) ------------------------------
    double nf[RPT], np[RPT];
    {
      // cell rows k-1 .. k+RPT-1 of the columns j-1 (L) and j (R); post_vol*density1 (:69-121)
      auto pm = [&](int c) {
        double post_vol;
        if (MS == 1) post_vol = svol[c] + svf[c + BW] - svf[c];
        else if (MS == 2) post_vol = svol[c] + svf[c + 1] - svf[c];
        else post_vol = svol[c];
        return sd1[c] * post_vol;
      };
      double pmL[RPT + 1], pmR[RPT + 1], m0[RPT + 1], m1[RPT + 1];
      // x sweep: plane column 0 (node j0-2) would need cell column j0-3, y sweep: plane row 0 (node k0-2) cell row
      // k0-3; neither node mass is ever used, the index is kept inside the box
      const int bc = b0 + ((DIR == 1 && lx == 0) ? 1 : 0);
#pragma unroll
      for (int i = 0; i <= RPT; ++i) {
        const int ro = (DIR == 2 && i == 0 && row0 == 0) ? 0 : (i - 1) * BW;
        pmR[i] = pm(bc + ro);
        pmL[i] = pm(bc + ro - 1);
        if (DIR == 1) {  // mass_flux_x(j, k-1+i), (j+1, k-1+i)
          m0[i] = smf[b0 + (i - 1) * BW];
          m1[i] = smf[b0 + (i - 1) * BW + 1];
        } else {         // mass_flux_y(j-1, k+i), (j, k+i)
          m0[i] = smf[b0 + i * BW - 1];
          m1[i] = smf[b0 + i * BW];
        }
      }
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        if (DIR == 1) nf[r] = 0.25 * (m0[r] + m0[r + 1] + m1[r] + m1[r + 1]);  // (j,k-1)+(j,k)+(j+1,k-1)+(j+1,k)
        else          nf[r] = 0.25 * (m0[r] + m1[r] + m0[r + 1] + m1[r + 1]);  // (j-1,k)+(j,k)+(j-1,k+1)+(j,k+1)
        np[r] = 0.25 * (pmR[r] + pmR[r + 1] + pmL[r] + pmL[r + 1]);            // (j,k-1)+(j,k)+(j-1,k-1)+(j-1,k)
        s_nf[p0 + r * TX] = nf[r];
        s_np[p0 + r * TX] = np[r];
      }
    }
    __syncthreads();
    if (sched) queue.step(stage);  // everybody has read this iteration's table slot
    // ---- B: mom_flux at sweep positions 1 .. NPS-3 ---------------------------------------------------------------------
    double ma[RPT], mb[RPT], va0[RPT], vb0[RPT], nm_pre[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int ps = DIR == 1 ? lx : row0 + r;
      const int p = p0 + r * TX, b = b0 + r * BW;
      const bool mine = ps >= 1 && ps <= NPS - 3;
      // out-of-range positions read inside the planes / the box and are never stored
      const int pm1 = ps >= 1 ? p - SP : p, pp1 = ps <= NPS - 2 ? p + SP : p;
      const int bm1 = ps >= 1 ? b - SB : b, bp1 = ps <= NPS - 2 ? b + SB : b, bp2 = ps <= NPS - 3 ? b + 2 * SB : b;
      const double f = nf[r], f_m = s_nf[pm1], f_p = s_nf[pp1];
      nm_pre[r] = np[r] - f_m + f;
      const double nm_pre_p = s_np[pp1] - f + f_p;
      const bool neg = f < 0.0;
      const double width = cw[r], width_dif = neg ? cwp[r] : cwm[r];
      const double nmp_don = neg ? nm_pre_p : nm_pre[r];
      const double a_m = sva[bm1], a_0 = sva[b], a_p = sva[bp1], a_pp = sva[bp2];
      const double b_m = svb[bm1], b_0 = svb[b], b_p = svb[bp1], b_pp = svb[bp2];
      va0[r] = a_0;
      vb0[r] = b_0;
      // (the zero-numerator and inactive-limiter short cuts inside mom_face_flux skip most of the divisions on the
      // quiescent part of the mesh; a branch-free evaluation was measured 1.5x slower on clover_bm16)
      ma[r] = mom_face_flux(f, nmp_don, neg ? a_pp : a_m, neg ? a_p : a_0, neg ? a_0 : a_p, width, width_dif);
      mb[r] = mom_face_flux(f, nmp_don, neg ? b_pp : b_m, neg ? b_p : b_0, neg ? b_0 : b_p, width, width_dif);
      if (mine) {
        s_ma[p] = ma[r];
        s_mb[p] = mb[r];
      }
    }
    __syncthreads();
    // ---- C: the velocity update at sweep positions 2 .. NPS-3 ------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int ps = DIR == 1 ? lx : row0 + r;
      const int j = DIR == 1 ? j0 - 2 + lx : j0 + lx;
      const int k = DIR == 1 ? k0 + row0 + r : k0 - 2 + row0 + r;
      if (ps >= 2 && ps <= NPS - 3 && j <= nx + 1 && k <= ny + 1) {
        const int p = p0 + r * TX;
        // the flux through the previous position: my own when the rows of a thread run along the sweep
        const double ma_m = (DIR == 2 && r > 0) ? ma[r > 0 ? r - 1 : 0] : s_ma[p - SP];
        const double mb_m = (DIR == 2 && r > 0) ? mb[r > 0 ? r - 1 : 0] : s_mb[p - SP];
        const size_t o = idx2(pitch, j, k);
        va_new[o] = ddiv(va0[r] * nm_pre[r] + ma_m - ma[r], np[r]);
        vb_new[o] = ddiv(vb0[r] * nm_pre[r] + mb_m - mb[r], np[r]);
      }
    }
    __syncthreads();  // stage and planes are free again
  }
  gate.finish();
  if (sched) queue.leave();
  if (dep_start != 0) {
    ring_copy(va_old, va_new, nx, ny, pitch, 1, (int)blockIdx.x * NT + tid, G * NT);
    ring_copy(vb_old, vb_new, nx, ny, pitch, 1, (int)blockIdx.x * NT + tid, G * NT);
  }
}

template <int DIR, int MS, int TX, int TY, int RPT, int STAGES, int CPS>
static void launch_mom(const Grid& g, const MomMaps& M, const double* va_old, double* va_new, const double* vb_old,
                       double* vb_new, const double* celld) {
  using Cfg = MomCfg<DIR, TX, TY, RPT, STAGES, CPS, MS>;
  static bool configured = false;
  if (!configured) {
    CLV_CUDA(cudaFuncSetAttribute(advec_mom_tma_kernel<DIR, MS, TX, TY, RPT, STAGES, CPS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  const int ntx = (g.nx + 1 + Cfg::W - 1) / Cfg::W, nty = (g.ny + 1 + Cfg::H - 1) / Cfg::H;
  const int ntiles = ntx * nty;
  const int cap = sm_count() * CPS;
  const int ctas = ntiles < cap ? ntiles : cap;
  const TileOrder ord = tile_order_split(ntx, nty, Cfg::W, Cfg::H, Cfg::OX, Cfg::BW - Cfg::OX - Cfg::W, Cfg::OY,
                                         Cfg::BH - Cfg::OY - Cfg::H, g.nx, g.ny);
  launch_pdl(advec_mom_tma_kernel<DIR, MS, TX, TY, RPT, STAGES, CPS>, dim3(ctas), dim3(Cfg::NT), Cfg::SMEM, stream(), M, va_old,
             va_new, vb_old, vb_new, celld, g.nx, g.ny, g.pitch, ntx, ntiles, ord.table, next_tickets(),
             dep_start_for(ord), current_trace());
}

// ====================================================================================================================
// advec_mom, y sweep, "column march" (see advec_cell_ymarch_tma_kernel below for the idea): thread (lx, grp) owns node
// column j0+lx and the MM_R nodes k0+grp*MM_R .. of a 32 x (8*MM_R) tile.  Node fluxes, node masses, the MM_R+1 limited
// momentum fluxes and the MM_R velocity updates of a thread all come from its own column (and column j-1) of the staged
// boxes; nothing travels between threads, so there are no planes and no barriers between phases.  The flux through the
// node face under a group is evaluated twice (once by the group below) instead of handed over behind a barrier.
// Boxes: columns j0-2 .. j0+33, rows k0-2 .. k0+H+1.  MM_R = 3 keeps the six-box ring at two CTAs per SM.
#ifndef MM_R_
#define MM_R_ 3
#endif
#ifndef MM_CPS
#define MM_CPS 2
#endif
#ifndef MM_G_
#define MM_G_ 8
#endif
constexpr int MM_R = MM_R_, MM_G = MM_G_, MM_W = 32, MM_H = MM_G * MM_R, MM_BW = MM_W + 4, MM_BH = MM_H + 4, MM_STAGES = 2;
template <int MS>
using MomMarchRing = TileRing<mom_narr(MS), MM_BW, MM_BH, MM_STAGES>;
template <int MS>
constexpr int mm_smem() { return MomMarchRing<MS>::BYTES + 128; }
static_assert(fits_sm(mm_smem<1>(), MM_CPS), "advec_mom march: MM_CPS CTAs do not fit one SM");

template <int MS>  // mom_sweep 2 (first sweep along y) or 4 (second sweep along y)
__global__ void __launch_bounds__(MM_W* MM_G, MM_CPS)
    advec_mom_ymarch_tma_kernel(const __grid_constant__ MomMaps M, const double* __restrict__ va_old, double* __restrict__ va_new,
                                const double* __restrict__ vb_old, double* __restrict__ vb_new,
                                const double* __restrict__ celld, int nx, int ny, int pitch, int ntiles,
                                const int2* __restrict__ order, Tickets tickets, int dep_start, unsigned long long* trace) {
  constexpr int NT = MM_W * MM_G, BW = MM_BW, R = MM_R;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = align128(smem_raw);
  MomMarchRing<MS> ring;
  ring.init(smem);
  const int tid = threadIdx.x, lx = tid % MM_W, grp = tid / MM_W;
  const int G = gridDim.x;
  pdl_trigger();
  PdlGate gate(dep_start, trace);
  if (dep_start == 0) {
    gate.need(0);
    ring_copy(va_old, va_new, nx, ny, pitch, 1, (int)blockIdx.x * NT + tid, G * NT);
    ring_copy(vb_old, vb_new, nx, ny, pitch, 1, (int)blockIdx.x * NT + tid, G * NT);
  }
  auto issue_tile = [&](int stage, int2 xy) {
    const int j0 = 1 + xy.x * MM_W, k0 = 1 + xy.y * MM_H;
    ring.issue(M.m, stage, j0 - 2 + XOFF, k0 - 2 + 1);
  };
  __shared__ int s_tile[4];  // four entries indexed by iteration (see advec_cell_ymarch_tma_kernel)
  __shared__ int2 s_xy[4];
  __shared__ int s_q[8];
  TileQueue<MM_STAGES> queue(tickets, ntiles, order, s_tile, s_xy, s_q);
  const bool sched = (tid == 32);
  if (sched) queue.prime_all();
  __syncthreads();
  if (tid == 0 && s_tile[0] < ntiles) {
    gate.need(s_tile[0]);
    issue_tile(0, s_xy[0]);
  }
  const int smax = ny + 2;
  for (int it = 0;; ++it) {
    const int stage = it % MM_STAGES;
    const int t = s_tile[it & 3];
    if (t >= ntiles) break;
    const int2 cur = s_xy[it & 3];
    gate.need(t);
    if (sched) queue.step((it + 2) & 3);
    if (tid == 0) {
      const int tn = s_tile[(it + 1) & 3];
      if (tn < ntiles) {
        gate.need(tn);
        issue_tile((stage + 1) % MM_STAGES, s_xy[(it + 1) & 3]);
      }
    }
    const int j0 = 1 + cur.x * MM_W, k0 = 1 + cur.y * MM_H;
    const int j = j0 + lx, kA = k0 + grp * R;
    // celldy at kA-2 .. kA+R (1-D, lower bound -1 -> index k+1; clamped for rows that are never used)
    double cd[R + 3];
#pragma unroll
    for (int i = 0; i < R + 3; ++i) cd[i] = celld[clampi(kA - 2 + i, -1, smax) + 1];
    ring.wait(stage, (uint32_t)((it / MM_STAGES) & 1));
    const double* __restrict__ svol = ring.tile(stage, MA_VOLUME);
    const double* __restrict__ sd1 = ring.tile(stage, MA_DENSITY1);
    const double* __restrict__ smf = ring.tile(stage, MA_MASS_FLUX);
    const double* __restrict__ sva = ring.tile(stage, MA_VEL_A);
    const double* __restrict__ svb = ring.tile(stage, MA_VEL_B);
    const double* __restrict__ svf = ring.tile(stage, MS <= 2 ? MA_VOL_FLUX : MA_VOLUME);  // (read for MS 2 only)
    if (j <= nx + 1 && kA <= ny + 1) {
      const int bA = (grp * R + 2) * BW + lx + 2;  // box position of (j, kA); (j, kA+i) at bA + i*BW
      auto pm = [&](int c) {                       // post_vol * density1 of a cell (:69-121)
        const double post_vol = (MS == 2) ? svol[c] + svf[c + 1] - svf[c] : svol[c];
        return sd1[c] * post_vol;
      };
      // node_mass_post of the nodes kA-1 .. kA+R (:216-231): cells (j,k-1)+(j,k)+(j-1,k-1)+(j-1,k)
      double np[R + 2];
      {
        double pmR[R + 3], pmL[R + 3];
#pragma unroll
        for (int i = 0; i < R + 3; ++i) {
          pmR[i] = pm(bA + (i - 2) * BW);
          pmL[i] = pm(bA + (i - 2) * BW - 1);
        }
#pragma unroll
        for (int i = 0; i < R + 2; ++i) np[i] = 0.25 * (pmR[i] + pmR[i + 1] + pmL[i] + pmL[i + 1]);
      }
      // node_flux of the nodes kA-2 .. kA+R (:205-214): mass_flux_y (j-1,k)+(j,k)+(j-1,k+1)+(j,k+1)
      double nf[R + 3];
      {
        double m0[R + 4], m1[R + 4];
#pragma unroll
        for (int i = 0; i < R + 4; ++i) {
          m0[i] = smf[bA + (i - 2) * BW - 1];
          m1[i] = smf[bA + (i - 2) * BW];
        }
#pragma unroll
        for (int i = 0; i < R + 3; ++i) nf[i] = 0.25 * (m0[i] + m1[i] + m0[i + 1] + m1[i + 1]);
      }
      // velocities at kA-2 .. kA+R+1
      double va[R + 4], vb[R + 4];
#pragma unroll
      for (int i = 0; i < R + 4; ++i) {
        va[i] = sva[bA + (i - 2) * BW];
        vb[i] = svb[bA + (i - 2) * BW];
      }
      // the momentum fluxes through the node faces kA-1 .. kA+R-1 (:240-270); face kf <-> nf[f+1], np[f], cd[f+1]
      double ma[R + 1], mb[R + 1], nmpre[R + 1];
#pragma unroll
      for (int f = 0; f <= R; ++f) {
        const double fl = nf[f + 1], f_m = nf[f], f_p = nf[f + 2];
        nmpre[f] = np[f] - f_m + fl;
        const double nm_pre_p = np[f + 1] - fl + f_p;
        const bool neg = fl < 0.0;
        const double width = cd[f + 1], width_dif = neg ? cd[f + 2] : cd[f];
        const double nmp_don = neg ? nm_pre_p : nmpre[f];
        ma[f] = 0.0;
        mb[f] = 0.0;
        if (kA - 1 + f <= ny + 1) {
          ma[f] = mom_face_flux(fl, nmp_don, neg ? va[f + 3] : va[f], neg ? va[f + 2] : va[f + 1], neg ? va[f + 1] : va[f + 2], width, width_dif);
          mb[f] = mom_face_flux(fl, nmp_don, neg ? vb[f + 3] : vb[f], neg ? vb[f + 2] : vb[f + 1], neg ? vb[f + 1] : vb[f + 2], width, width_dif);
        }
      }
      // the nodes kA .. kA+R-1 (:272-282)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int k = kA + r;
        if (k <= ny + 1) {
          const size_t o = idx2(pitch, j, k);
          va_new[o] = ddiv(va[r + 2] * nmpre[r + 1] + ma[r] - ma[r + 1], np[r + 1]);
          vb_new[o] = ddiv(vb[r + 2] * nmpre[r + 1] + mb[r] - mb[r + 1], np[r + 1]);
        }
      }
    }
    __syncthreads();  // the stage is free again
  }
  gate.finish();
  if (sched) queue.leave();
  if (dep_start != 0) {
    ring_copy(va_old, va_new, nx, ny, pitch, 1, (int)blockIdx.x * NT + tid, G * NT);
    ring_copy(vb_old, vb_new, nx, ny, pitch, 1, (int)blockIdx.x * NT + tid, G * NT);
  }
}

template <int MS>
static void launch_mom_ymarch(const Grid& g, const MomMaps& M, const double* va_old, double* va_new, const double* vb_old,
                              double* vb_new, const double* celld) {
  static bool configured = false;
  if (!configured) {
    CLV_CUDA(cudaFuncSetAttribute(advec_mom_ymarch_tma_kernel<MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, mm_smem<MS>()));
    configured = true;
  }
  const int ntx = (g.nx + 1 + MM_W - 1) / MM_W, nty = (g.ny + 1 + MM_H - 1) / MM_H;
  const int ntiles = ntx * nty;
  const int cap = sm_count() * MM_CPS;
  const int ctas = ntiles < cap ? ntiles : cap;
  const TileOrder ord = tile_order_split(ntx, nty, MM_W, MM_H, 2, MM_BW - 2 - MM_W, 2, MM_BH - 2 - MM_H, g.nx, g.ny);
  launch_pdl(advec_mom_ymarch_tma_kernel<MS>, dim3(ctas), dim3(MM_W * MM_G), mm_smem<MS>(), stream(), M, va_old, va_new, vb_old, vb_new,
             celld, g.nx, g.ny, g.pitch, ntiles, ord.table, next_tickets(), dep_start_for(ord), current_trace());
}
static bool mom_ymarch_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("CLOVER_B200_MOM_YMARCH") ? atoi(getenv("CLOVER_B200_MOM_YMARCH")) : 1;  // 0: the three-phase kernel (A/B)
  return v != 0;
}

// ====================================================================================================================
// advec_cell (advec_cell_kernel_c.c:72-177 x sweep, :182-290 y sweep).  `s` is the sweep axis, the tile owns the cells
// s0 .. s0+NS-1 and the thread grid has N = NS+3(+1) positions along the sweep, position i <-> index s0-1+i:
//   A  pre_vol(s)                   all positions                  (:72-102 / :182-214)
//   B  mass_flux(s), ener_flux(s)   positions 1 .. N-1  (face s = lower face of cell s)   (:107-151 / :219-263)
//   C  density1(s), energy1(s)      positions 1 .. NS                                      (:156-175 / :266-286)
// mass_flux is stored for the faces of the tile's own cells; the last tile along the sweep also stores the faces
// n+1 and n+2 (the loop of the reference runs to x_max+2 / y_max+2).
// Boxes: CA_VF is the volume flux along the sweep, CA_VFC the one across it.  Only the pre-volume of the first sweep of a
// step reads the cross flux (:77-81 / :189-193 against :94-96 / :207-209), so the second sweep neither loads that box
// nor gives it room (one of eight passes less).
enum { CA_VOLUME = 0, CA_VF, CA_DENSITY1, CA_ENERGY1, CA_VFC, CA_NARR };
constexpr int cell_narr(int sweep) { return sweep == 1 ? CA_NARR : CA_NARR - 1; }

template <int DIR, int TX, int TY, int RPT, int STAGES, int CPS, int SWEEP = 1>
struct CellCfg {
  static constexpr int NT = TX * TY;
  static constexpr int ROWS = TY * RPT;
  static constexpr int W = DIR == 1 ? TX - 4 : TX;      // cells per tile along x (even: TMA alignment)
  static constexpr int H = DIR == 1 ? ROWS : ROWS - 3;  // cells per tile along y
  static constexpr int BW = DIR == 1 ? TX + 2 : TX + 4; // box: x from j0-2
  static constexpr int BH = DIR == 1 ? ROWS + 1 : ROWS + 2;  // box: y from k0 (x sweep) / k0-2 (y sweep)
  static constexpr int OX = 2, OY = DIR == 1 ? 0 : 2;
  using Ring = TileRing<cell_narr(SWEEP), BW, BH, STAGES>;
  static constexpr int NI = TX * ROWS;
  // planes of their own: pre_vol, and in the second sweep the energy flux; the mass flux lives in the volume box and the
  // first sweep's energy flux in the cross-flux box, both dead once phase A is behind its barrier
  static_assert(NI <= BW * BH, "a plane must fit a box");
  static constexpr int NPLANES = SWEEP == 1 ? 1 : 2;
  static constexpr int SMEM = Ring::BYTES + NPLANES * NI * 8 + 128;
  static_assert(fits_sm(SMEM, CPS), "advec_cell: CPS CTAs of this shape do not fit one SM");
};
struct CellMaps {
  CUtensorMap m[CA_NARR];
};

template <int DIR, int SWEEP, int TX, int TY, int RPT, int STAGES, int CPS>
__global__ void __launch_bounds__(TX* TY, CPS)
    advec_cell_tma_kernel(const __grid_constant__ CellMaps M, const double* __restrict__ d_old, double* __restrict__ d_new,
                          const double* __restrict__ e_old, double* __restrict__ e_new,
                          double* __restrict__ mass_flux, const double* __restrict__ vertexd, int nx, int ny, int pitch,
                          int ntx, int nty, const int2* __restrict__ order, Tickets tickets, int dep_start,
                          unsigned long long* trace) {
  using Cfg = CellCfg<DIR, TX, TY, RPT, STAGES, CPS, SWEEP>;
  constexpr int NT = Cfg::NT, W = Cfg::W, H = Cfg::H, BW = Cfg::BW, NI = Cfg::NI, OX = Cfg::OX, OY = Cfg::OY;
  constexpr int NS = DIR == 1 ? W : H;             // cells of a tile along the sweep
  constexpr int NPS = DIR == 1 ? TX : Cfg::ROWS;   // plane positions along the sweep
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = align128(smem_raw);
  typename Cfg::Ring ring;
  ring.init(smem);
  double* __restrict__ s_pv = reinterpret_cast<double*>(smem + Cfg::Ring::BYTES);  // pre_vol
  const int tid = threadIdx.x, lx = tid % TX, ty = tid / TX;
  const int G = gridDim.x;
  const int ntiles = ntx * nty;
  pdl_trigger();
  PdlGate gate(dep_start, trace);
  // The halo ring of the old buffers (what the preceding halo exchange / reflective boundary delivered) moves to the
  // new ones: up front, next to the first TMA loads, when this launch waits for its predecessor anyway; after the
  // tiles when the interior tiles run ahead of a halo kernel (dep_start > 0).
  if (dep_start == 0) {
    gate.need(0);
    ring_copy(d_old, d_new, nx, ny, pitch, 0, (int)blockIdx.x * NT + tid, G * NT);
    ring_copy(e_old, e_new, nx, ny, pitch, 0, (int)blockIdx.x * NT + tid, G * NT);
  }
  auto issue_tile = [&](int stage, int2 xy) {
    const int j0 = 1 + xy.x * W, k0 = 1 + xy.y * H;
    ring.issue(M.m, stage, j0 - OX + XOFF, k0 - OY + 1);
  };
  __shared__ int s_tile[STAGES];
  __shared__ int2 s_xy[STAGES];
  __shared__ int s_q[8];
  TileQueue<STAGES> queue(tickets, ntiles, order, s_tile, s_xy, s_q);
  const bool sched = (tid == 32);  // lane 0 of warp 1 drives the tile queue (tma.cuh)
  if (sched) queue.prime_all();
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s_tile[s] < ntiles) {
        gate.need(s_tile[s]);
        issue_tile(s, s_xy[s]);
      }
    }
  }
  constexpr int SB = DIR == 1 ? 1 : BW;  // box stride along the sweep
  constexpr int SC = DIR == 1 ? BW : 1;  // box stride across the sweep
  constexpr int SP = DIR == 1 ? 1 : TX;  // plane stride along the sweep
  const int smax = (DIR == 1 ? nx : ny) + 2;
  for (int it = 0;; ++it) {
    const int stage = it % STAGES;
    const int t = s_tile[stage];
    if (t >= ntiles) break;
    const int2 cur = s_xy[stage];
    gate.need(t);
    if (tid == 0) {
      const int ns = (stage + STAGES - 1) % STAGES;
      const int tn = s_tile[ns];
      if (tn < ntiles) {
        gate.need(tn);
        issue_tile(ns, s_xy[ns]);
      }
    }
    const int tx_ = cur.x, ty_ = cur.y;
    const int j0 = 1 + tx_ * W, k0 = 1 + ty_ * H;
    const bool last_along = DIR == 1 ? (tx_ == ntx - 1) : (ty_ == nty - 1);
    const int row0 = ty * RPT;
    // vertexdx / vertexdy at s, s-1 and min(s+1, n+2) (1-D, lower bound -1 -> index s+1); before the wait
    double vd0[RPT], vdm[RPT], vdu[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int s_face = DIR == 1 ? j0 - 1 + lx : k0 - 1 + row0 + r;
      const int sup = s_face + 1 < smax ? s_face + 1 : smax;  // MIN(j+1, x_max+2), :114
      vd0[r] = vertexd[clampi(s_face, -1, smax) + 1];
      vdm[r] = vertexd[clampi(s_face - 1, -1, smax) + 1];
      vdu[r] = vertexd[clampi(sup, -1, smax) + 1];
    }
    ring.wait(stage, (uint32_t)((it / STAGES) & 1));
    const double* svol = ring.tile(stage, CA_VOLUME);
    const double* __restrict__ svf = ring.tile(stage, CA_VF);        // the volume flux along the sweep
    const double* svc = ring.tile(stage, SWEEP == 1 ? CA_VFC : CA_VOLUME);  // across it (read for SWEEP 1 only)
    double* s_mf = ring.scratch(stage, CA_VOLUME);                   // mass flux through the lower face   (phases B, C)
    double* s_ef = SWEEP == 1 ? ring.scratch(stage, CA_VFC) : s_pv + NI;  // energy flux                  (phases B, C)
    const double* __restrict__ sd = ring.tile(stage, CA_DENSITY1);
    const double* __restrict__ se = ring.tile(stage, CA_ENERGY1);
    // x sweep: plane column lx <-> cell j0-1+lx (box column lx+1), plane row <-> k0+row (box row row).
    // y sweep: plane column lx <-> cell j0+lx (box column lx+2),   plane row <-> k0-1+row (box row row+1).
    const int b0 = DIR == 1 ? row0 * BW + lx + 1 : (row0 + 1) * BW + lx + OX;
    const int p0 = row0 * TX + lx;
    // ---- A: pre_vol ------------------------------------------------------------------------------------------------
    double pv[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int c = b0 + r * BW;
      // the flux difference along the sweep first, then the one across it: :77-81 / :189-193; sweep 2: :94-96 / :207-209
      if (SWEEP == 1) pv[r] = svol[c] + (svf[c + SB] - svf[c] + svc[c + SC] - svc[c]);
      else            pv[r] = svol[c] + svf[c + SB] - svf[c];
      s_pv[p0 + r * TX] = pv[r];
    }
    __syncthreads();
    if (sched) queue.step(stage);  // everybody has read this iteration's table slot
    // ---- B: fluxes through the lower face of position ps (valid for ps >= 1) --------------------------------------------
    double mf[RPT], ef[RPT], vf[RPT], d0[RPT], e0[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int ps = DIR == 1 ? lx : row0 + r;
      const int p = p0 + r * TX, b = b0 + r * BW;
      const int s_face = DIR == 1 ? j0 - 1 + lx : k0 - 1 + row0 + r;
      const int bm1 = b - SB, bm2 = ps >= 1 ? b - 2 * SB : b - SB;   // position 0 never stores
      const int bup = (s_face + 1 <= smax) ? b + SB : b;            // MIN(j+1, x_max+2)
      vf[r] = svf[b];
      const double pv_m = s_pv[ps >= 1 ? p - SP : p];
      const double dm2 = sd[bm2], dm1 = sd[bm1], dp1 = sd[bup];
      const double em2 = se[bm2], em1 = se[bm1], ep1 = se[bup];
      d0[r] = sd[b];
      e0[r] = se[b];
      const bool pos = vf[r] > 0.0;
      const double pvd = pos ? pv_m : pv[r];
      const double vdd = pos ? vdm[r] : vdu[r];
      cell_face_flux(vf[r], pvd, pos ? dm2 : dp1, pos ? dm1 : d0[r], pos ? d0[r] : dm1, pos ? em2 : ep1,
                     pos ? em1 : e0[r], pos ? e0[r] : em1, vd0[r], vdd, mf[r], ef[r]);
      s_mf[p] = mf[r];
      s_ef[p] = ef[r];
      const int j = DIR == 1 ? j0 - 1 + lx : j0 + lx;
      const int k = DIR == 1 ? k0 + row0 + r : k0 - 1 + row0 + r;
      const bool cross_ok = DIR == 1 ? (k <= ny) : (j <= nx);
      if (ps >= 1 && (ps <= NS || last_along) && s_face <= smax && cross_ok) mass_flux[idx2(pitch, j, k)] = mf[r];
    }
    __syncthreads();
    // ---- C: the cell update at positions 1 .. NS -----------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int ps = DIR == 1 ? lx : row0 + r;
      const int j = DIR == 1 ? j0 - 1 + lx : j0 + lx;
      const int k = DIR == 1 ? k0 + row0 + r : k0 - 1 + row0 + r;
      if (ps >= 1 && ps <= NS && j <= nx && k <= ny) {
        const int p = p0 + r * TX, b = b0 + r * BW;
        const bool own_next = DIR == 2 && r + 1 < RPT;  // the next face along the sweep is my own next row
        const double mf_p = own_next ? mf[r + 1 < RPT ? r + 1 : r] : s_mf[p + SP];
        const double ef_p = own_next ? ef[r + 1 < RPT ? r + 1 : r] : s_ef[p + SP];
        const double vf_p = own_next ? vf[r + 1 < RPT ? r + 1 : r] : svf[b + SB];
        const double pre_mass = d0[r] * pv[r];
        const double post_mass = pre_mass + mf[r] - mf_p;
        const double post_ener = (e0[r] * pre_mass + ef[r] - ef_p) / post_mass;
        const double advec_vol = pv[r] + vf[r] - vf_p;
        const size_t o = idx2(pitch, j, k);
        d_new[o] = post_mass / advec_vol;
        e_new[o] = post_ener;
      }
    }
    __syncthreads();  // stage and planes are free again
  }
  gate.finish();
  if (sched) queue.leave();
  if (dep_start != 0) {
    ring_copy(d_old, d_new, nx, ny, pitch, 0, (int)blockIdx.x * NT + tid, G * NT);
    ring_copy(e_old, e_new, nx, ny, pitch, 0, (int)blockIdx.x * NT + tid, G * NT);
  }
}

template <int DIR, int SWEEP, int TX, int TY, int RPT, int STAGES, int CPS>
static void launch_cell(const Grid& g, const CellMaps& M, const double* d_old, double* d_new, const double* e_old,
                        double* e_new, double* mass_flux, const double* vertexd) {
  using Cfg = CellCfg<DIR, TX, TY, RPT, STAGES, CPS, SWEEP>;
  static bool configured = false;
  if (!configured) {
    CLV_CUDA(cudaFuncSetAttribute(advec_cell_tma_kernel<DIR, SWEEP, TX, TY, RPT, STAGES, CPS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  const int ntx = (g.nx + Cfg::W - 1) / Cfg::W, nty = (g.ny + Cfg::H - 1) / Cfg::H;
  const int ntiles = ntx * nty;
  const int cap = sm_count() * CPS;
  const int ctas = ntiles < cap ? ntiles : cap;
  const TileOrder ord = tile_order_split(ntx, nty, Cfg::W, Cfg::H, Cfg::OX, Cfg::BW - Cfg::OX - Cfg::W, Cfg::OY,
                                         Cfg::BH - Cfg::OY - Cfg::H, g.nx, g.ny);
  launch_pdl(advec_cell_tma_kernel<DIR, SWEEP, TX, TY, RPT, STAGES, CPS>, dim3(ctas), dim3(Cfg::NT), Cfg::SMEM, stream(), M, d_old,
             d_new, e_old, e_new, mass_flux, vertexd, g.nx, g.ny, g.pitch, ntx, nty, ord.table, next_tickets(),
             dep_start_for(ord), current_trace());
}


// ====================================================================================================================
// advec_cell, y sweep, "column march": the same arithmetic with NO intermediate planes and NO barriers between phases.
// Thread (lx, grp) owns column j0+lx and the YM_R cells k0+grp*YM_R .. +YM_R-1 of a 32 x (8*YM_R) tile.  Everything a
// face or a cell needs along y comes from the thread's own column of the staged boxes (pre_vol needs vol_flux_x at j and
// j+1, which are inputs, not results), so the YM_R+1 face fluxes and YM_R cell updates of a thread are independent
// straight-line work; the face on top of a group is evaluated twice (by the group above as its own) instead of passed
// through shared memory behind a barrier.  Boxes: columns j0-2 .. j0+33, rows k0-2 .. k0+H+2.  The tiling covers the
// rows 1 .. ny+2: the reference stores mass_flux_y up to face y_max+2 (advec_cell_kernel_c.c:219), cells beyond ny are
// not updated.  ncu on the three-phase kernel: top stalls `wait` and `barrier`, 23 % of the rows of a box are halo.
#ifndef YM_R_
#define YM_R_ 4
#endif
#ifndef YM_CPS
#define YM_CPS 2
#endif
#ifndef YM_G_
#define YM_G_ 8
#endif
constexpr int YM_R = YM_R_, YM_G = YM_G_, YM_W = 32, YM_H = YM_G * YM_R, YM_BW = YM_W + 4, YM_BH = YM_H + 5, YM_STAGES = 2;
template <int SWEEP>
using MarchRing = TileRing<cell_narr(SWEEP), YM_BW, YM_BH, YM_STAGES>;
template <int SWEEP>
constexpr int ym_smem() { return MarchRing<SWEEP>::BYTES + 128; }
static_assert(fits_sm(ym_smem<1>(), YM_CPS), "advec_cell march: YM_CPS CTAs do not fit one SM");

template <int SWEEP>
__global__ void __launch_bounds__(YM_W* YM_G, YM_CPS)
    advec_cell_ymarch_tma_kernel(const __grid_constant__ CellMaps M, const double* __restrict__ d_old, double* __restrict__ d_new,
                                 const double* __restrict__ e_old, double* __restrict__ e_new,
                                 double* __restrict__ mass_flux, const double* __restrict__ vertexd, int nx, int ny, int pitch,
                                 int ntiles, const int2* __restrict__ order, Tickets tickets, int dep_start,
                                 unsigned long long* trace) {
  constexpr int NT = YM_W * YM_G, BW = YM_BW;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = align128(smem_raw);
  MarchRing<SWEEP> ring;
  ring.init(smem);
  const int tid = threadIdx.x, lx = tid % YM_W, grp = tid / YM_W;
  const int G = gridDim.x;
  pdl_trigger();
  PdlGate gate(dep_start, trace);
  if (dep_start == 0) {  // (as in advec_cell_tma_kernel: the halo ring of the old buffers moves to the new ones)
    gate.need(0);
    ring_copy(d_old, d_new, nx, ny, pitch, 0, (int)blockIdx.x * NT + tid, G * NT);
    ring_copy(e_old, e_new, nx, ny, pitch, 0, (int)blockIdx.x * NT + tid, G * NT);
  }
  auto issue_tile = [&](int stage, int2 xy) {
    const int j0 = 1 + xy.x * YM_W, k0 = 1 + xy.y * YM_H;
    ring.issue(M.m, stage, j0 - 2 + XOFF, k0 - 2 + 1);
  };
  // Tile table with FOUR entries, indexed by iteration: entry it%4 = the tile computed at iteration it.  With one barrier
  // per iteration the scheduler can then run two iterations ahead of the readers (it writes entry (it+2)%4 while the
  // leader reads (it+1)%4 and everybody reads it%4).
  __shared__ int s_tile[4];
  __shared__ int2 s_xy[4];
  __shared__ int s_q[8];
  TileQueue<YM_STAGES> queue(tickets, ntiles, order, s_tile, s_xy, s_q);
  const bool sched = (tid == 32);
  if (sched) queue.prime_all();  // entries 0 and 1
  __syncthreads();
  if (tid == 0 && s_tile[0] < ntiles) {
    gate.need(s_tile[0]);
    issue_tile(0, s_xy[0]);
  }
  const int smax = ny + 2;
  for (int it = 0;; ++it) {
    const int stage = it % YM_STAGES;
    const int t = s_tile[it & 3];
    if (t >= ntiles) break;
    const int2 cur = s_xy[it & 3];
    gate.need(t);
    if (sched) queue.step((it + 2) & 3);
    if (tid == 0) {  // the other stage was released by the barrier that ended iteration it-1
      const int tn = s_tile[(it + 1) & 3];
      if (tn < ntiles) {
        gate.need(tn);
        issue_tile((stage + 1) % YM_STAGES, s_xy[(it + 1) & 3]);
      }
    }
    const int j0 = 1 + cur.x * YM_W, k0 = 1 + cur.y * YM_H;
    const int j = j0 + lx, kA = k0 + grp * YM_R;
    // vertexdy at kA-1 .. kA+5 (1-D, lower bound -1 -> index k+1; clamped for rows that are never used)
    double vd[YM_R + 3];
#pragma unroll
    for (int i = 0; i < YM_R + 3; ++i) vd[i] = vertexd[clampi(kA - 1 + i, -1, smax) + 1];
    ring.wait(stage, (uint32_t)((it / YM_STAGES) & 1));
    const double* __restrict__ svol = ring.tile(stage, CA_VOLUME);
    const double* __restrict__ sfy = ring.tile(stage, CA_VF);
    const double* __restrict__ sfx = ring.tile(stage, SWEEP == 1 ? CA_VFC : CA_VOLUME);  // (read for SWEEP 1 only)
    const double* __restrict__ sd = ring.tile(stage, CA_DENSITY1);
    const double* __restrict__ se = ring.tile(stage, CA_ENERGY1);
    if (j <= nx && kA <= smax) {
      const int b0 = (grp * YM_R + 2) * BW + lx + 2;  // box position of cell (j, kA); cell (j, kA+i) at b0 + i*BW
      // pre_vol of the cells kA-1 .. kA+YM_R (:189-193 / :207-209)
      double pv[YM_R + 2];
#pragma unroll
      for (int i = 0; i < YM_R + 2; ++i) {
        const int c = b0 + (i - 1) * BW;
        if (SWEEP == 1) pv[i] = svol[c] + (sfy[c + BW] - sfy[c] + sfx[c + 1] - sfx[c]);
        else            pv[i] = svol[c] + sfy[c + BW] - sfy[c];
      }
      // density / energy of the cells kA-2 .. kA+YM_R+1
      double d[YM_R + 4], e[YM_R + 4];
#pragma unroll
      for (int i = 0; i < YM_R + 4; ++i) {
        d[i] = sd[b0 + (i - 2) * BW];
        e[i] = se[b0 + (i - 2) * BW];
      }
      // the faces kA .. kA+YM_R (face k = lower face of cell k): :219-263
      double mf[YM_R + 1], ef[YM_R + 1], vf[YM_R + 1];
#pragma unroll
      for (int f = 0; f <= YM_R; ++f) {
        const int kf = kA + f;
        mf[f] = 0.0; ef[f] = 0.0; vf[f] = 0.0;
        if (kf <= smax) {
          vf[f] = sfy[b0 + f * BW];
          const bool pos = vf[f] > 0.0;
          const bool up_ok = kf + 1 <= smax;  // MIN(k+1, y_max+2)
          // cell kf is d[f+2]: upwind kf-2 | min(kf+1, ny+2), donor kf-1 | kf, downwind kf | kf-1
          const double d_up = pos ? d[f] : (up_ok ? d[f + 3] : d[f + 2]);
          const double e_up = pos ? e[f] : (up_ok ? e[f + 3] : e[f + 2]);
          const double vdd = pos ? vd[f] : (up_ok ? vd[f + 2] : vd[f + 1]);  // vd[i] = vertexdy(kA-1+i)
          cell_face_flux(vf[f], pos ? pv[f] : pv[f + 1], d_up, pos ? d[f + 1] : d[f + 2], pos ? d[f + 2] : d[f + 1], e_up,
                         pos ? e[f + 1] : e[f + 2], pos ? e[f + 2] : e[f + 1], vd[f + 1], vdd, mf[f], ef[f]);
          if (f < YM_R) mass_flux[idx2(pitch, j, kf)] = mf[f];  // own faces; the face on top belongs to the group above
        }
      }
      // the cells kA .. kA+YM_R-1: :266-286
#pragma unroll
      for (int r = 0; r < YM_R; ++r) {
        const int k = kA + r;
        if (k <= ny) {
          const double pre_mass = d[r + 2] * pv[r + 1];
          const double post_mass = pre_mass + mf[r] - mf[r + 1];
          const double post_ener = (e[r + 2] * pre_mass + ef[r] - ef[r + 1]) / post_mass;
          const double advec_vol = pv[r + 1] + vf[r] - vf[r + 1];
          const size_t o = idx2(pitch, j, k);
          d_new[o] = post_mass / advec_vol;
          e_new[o] = post_ener;
        }
      }
    }
    __syncthreads();  // the stage is free again
  }
  gate.finish();
  if (sched) queue.leave();
  if (dep_start != 0) {
    ring_copy(d_old, d_new, nx, ny, pitch, 0, (int)blockIdx.x * NT + tid, G * NT);
    ring_copy(e_old, e_new, nx, ny, pitch, 0, (int)blockIdx.x * NT + tid, G * NT);
  }
}

template <int SWEEP>
static void launch_cell_ymarch(const Grid& g, const CellMaps& M, const double* d_old, double* d_new, const double* e_old,
                               double* e_new, double* mass_flux, const double* vertexd) {
  static bool configured = false;
  if (!configured) {
    CLV_CUDA(cudaFuncSetAttribute(advec_cell_ymarch_tma_kernel<SWEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ym_smem<SWEEP>()));
    configured = true;
  }
  const int ntx = (g.nx + YM_W - 1) / YM_W, nty = (g.ny + 2 + YM_H - 1) / YM_H;  // rows 1 .. ny+2 (faces up to y_max+2)
  const int ntiles = ntx * nty;
  const int cap = sm_count() * YM_CPS;
  const int ctas = ntiles < cap ? ntiles : cap;
  const TileOrder ord = tile_order_split(ntx, nty, YM_W, YM_H, 2, YM_BW - 2 - YM_W, 2, YM_BH - 2 - YM_H, g.nx, g.ny);
  launch_pdl(advec_cell_ymarch_tma_kernel<SWEEP>, dim3(ctas), dim3(YM_W * YM_G), ym_smem<SWEEP>(), stream(), M, d_old, d_new, e_old, e_new,
             mass_flux, vertexd, g.nx, g.ny, g.pitch, ntiles, ord.table, next_tickets(), dep_start_for(ord), current_trace());
}
static bool ymarch_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("CLOVER_B200_YMARCH") ? atoi(getenv("CLOVER_B200_YMARCH")) : 1;  // 0: the three-phase kernel (A/B)
  return v != 0;
}

void run_advec_cell_tma(const Grid& g, int dir, int sweep, double* vertexdx, double* vertexdy, double* volume,
                        double* density1, double* energy1, double* mass_flux_x, double* vol_flux_x, double* mass_flux_y,
                        double* vol_flux_y) {
  const double* vol = dev(g, volume, CELL, IN);
  const double* fx = dev(g, vol_flux_x, XFACE, IN);
  const double* fy = dev(g, vol_flux_y, YFACE, IN);
  const double* d_old = dev(g, density1, CELL, INOUT);
  const double* e_old = dev(g, energy1, CELL, INOUT);
  double* d_new = dev_alt(g, density1, CELL);
  double* e_new = dev_alt(g, energy1, CELL);
  const double* vd = dir == 1 ? dev(g, vertexdx, X1D_VERT, IN) : dev(g, vertexdy, Y1D_VERT, IN);
  double* mf = dir == 1 ? dev(g, mass_flux_x, XFACE, OUT_FULL) : dev(g, mass_flux_y, YFACE, OUT_FULL);
  const double* in[CA_NARR] = {vol, dir == 1 ? fx : fy, d_old, e_old, dir == 1 ? fy : fx};
  CellMaps M;
  LaunchScope ls(dir == 1 ? "advec_cell_x_tma" : "advec_cell_y_tma");
#define CLV_CELL(DIR, TX, TY, RPT, ST, CPS)                                                                \
  do {                                                                                                     \
    using Cfg = CellCfg<DIR, TX, TY, RPT, ST, CPS>;                                                        \
    for (int a = 0; a < CA_NARR; ++a) M.m[a] = *tensor_map_for(g, in[a], Cfg::BW, Cfg::BH);                \
    if (sweep == 1) launch_cell<DIR, 1, TX, TY, RPT, ST, CPS>(g, M, d_old, d_new, e_old, e_new, mf, vd);   \
    else            launch_cell<DIR, 2, TX, TY, RPT, ST, CPS>(g, M, d_old, d_new, e_old, e_new, mf, vd);   \
  } while (0)
  // <thread grid TX x TY, rows per thread, ring stages, CTAs per SM>; measured on B200 at 3840^2:
  //   x: <64,4,2,2,3> 0.181 ms, <64,4,2,2,2> 0.221, <64,4,2,3,2> 0.224, <64,4,1,2,4> 0.236, <64,8,1,2,2> 0.236
  //   y: <32,8,2,2,3> 0.211 ms, <32,8,3,2,2> 0.217, <32,8,2,2,2> 0.247, <64,4,3,2,2> 0.245, <32,8,4,2,2> 0.342
  if (dir == 1) {
    // (a barrier-free x variant -- a warp per row segment, the fluxes of face j+1 by warp shuffle, 30-wide tiles -- was
    // bit-identical but slower: 0.187 vs 0.178 ms, profiles/r02_experiment_xrow.json)
    CLV_CELL(1, 64, CELLX_TY, CELLX_RPT, 2, CELLX_CPS);
  } else if (ymarch_enabled()) {
    for (int a = 0; a < CA_NARR; ++a) M.m[a] = *tensor_map_for(g, in[a], YM_BW, YM_BH);
    if (sweep == 1) launch_cell_ymarch<1>(g, M, d_old, d_new, e_old, e_new, mf, vd);
    else            launch_cell_ymarch<2>(g, M, d_old, d_new, e_old, e_new, mf, vd);
  } else {
    CLV_CELL(2, 32, 8, 2, 2, 3);
  }
#undef CLV_CELL
  swap_alt(density1);
  swap_alt(energy1);
}

// Both velocity components of one advec_mom sweep (advec_mom_driver.f90:85,108).
void run_advec_mom_tma(const Grid& g, int dirn, int sweep, double* vel_a, double* vel_b, double* mass_flux_x,
                       double* vol_flux_x, double* mass_flux_y, double* vol_flux_y, double* volume, double* density1,
                       double* celldx, double* celldy) {
  const int mom_sweep = dirn + 2 * (sweep - 1);
  const double* vol = dev(g, volume, CELL, IN);
  const double* d1 = dev(g, density1, CELL, IN);
  const double* fx = dev(g, vol_flux_x, XFACE, IN);
  const double* fy = dev(g, vol_flux_y, YFACE, IN);
  const double* va_old = dev(g, vel_a, VERTEX, INOUT);
  double* va_new = dev_alt(g, vel_a, VERTEX);
  const double* vb_old = dev(g, vel_b, VERTEX, INOUT);
  double* vb_new = dev_alt(g, vel_b, VERTEX);
  const double* mf = dirn == 1 ? dev(g, mass_flux_x, XFACE, IN) : dev(g, mass_flux_y, YFACE, IN);
  const double* cd = dirn == 1 ? dev(g, celldx, X1D_CELL, IN) : dev(g, celldy, Y1D_CELL, IN);
  // the post-volume of mom_sweep 1 needs vol_flux_y, of mom_sweep 2 vol_flux_x, of 3 and 4 neither (:69-121);
  // the sixth box of sweeps 3/4 is neither mapped into the ring nor loaded (mom_narr)
  const double* in[MA_NARR] = {vol, d1, mf, va_old, vb_old, mom_sweep == 1 ? fy : fx};
  MomMaps M;
  LaunchScope ls(dirn == 1 ? "advec_mom_x_tma" : "advec_mom_y_tma");
#define CLV_MOM(DIR, TX, TY, RPT, ST, CPS)                                                             \
  do {                                                                                                 \
    using Cfg = MomCfg<DIR, TX, TY, RPT, ST, CPS>;                                                     \
    for (int a = 0; a < MA_NARR; ++a) M.m[a] = *tensor_map_for(g, in[a], Cfg::BW, Cfg::BH);            \
    if (mom_sweep <= 2) launch_mom<DIR, DIR, TX, TY, RPT, ST, CPS>(g, M, va_old, va_new, vb_old, vb_new, cd);     \
    else                launch_mom<DIR, DIR + 2, TX, TY, RPT, ST, CPS>(g, M, va_old, va_new, vb_old, vb_new, cd); \
  } while (0)
  // <thread grid TX x TY, rows per thread, ring stages, CTAs per SM>; measured on B200 at 3840^2:
  //   x: <64,4,2,2,3> 0.177 ms (fits since the flux planes moved into dead boxes; 80 registers, no spills),
  //      <64,4,2,2,2> 0.204, <64,4,1,2,4> 0.216, <64,8,1,2,2> 0.221, <64,4,2,3,2> 0.229, <64,2,4,2,2> 0.305
  //   y: <32,8,3,2,2> 0.214 ms, <32,8,2,2,3> 0.236, <32,4,4,2,3> 0.240, <32,8,2,2,2> 0.247, <32,16,1,2,2> 0.255
  if (dirn == 1) {
    CLV_MOM(1, 64, MOMX_TY, MOMX_RPT, 2, MOMX_CPS);
  } else if (mom_ymarch_enabled()) {
    for (int a = 0; a < MA_NARR; ++a) M.m[a] = *tensor_map_for(g, in[a], MM_BW, MM_BH);
    if (mom_sweep == 2) launch_mom_ymarch<2>(g, M, va_old, va_new, vb_old, vb_new, cd);
    else                launch_mom_ymarch<4>(g, M, va_old, va_new, vb_old, vb_new, cd);
  } else {
    CLV_MOM(2, 32, 8, 3, 2, 2);
  }
#undef CLV_MOM
  swap_alt(vel_a);
  swap_alt(vel_b);
}

}  // namespace clv
