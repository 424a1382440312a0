// fuse.cu -- several reference calls in ONE kernel (resident mode, see "deferred execution" in common.cuh).
//
// runtime.cu hands fuse_at() the recorded calls of one stretch of the hydro step.  The patterns below are the
// runs of calls that the reference driver always makes back to back (hydro.f90:52-64, timestep.f90:56-117,
// PdV.f90:46-138, advection.f90:43-110) and whose intermediates no other call reads:
//
//   T  ideal_gas(d0,e0) -> [exchange][update_halo]{d0,e0,p,u0,v0; depth 1} -> viscosity
//        -> [exchange][update_halo]{q; depth 1} -> calc_dt
//      one kernel: pressure of the four neighbours is recomputed from their density/energy (2 multiplies,
//      bit-identical to the exchanged/reflected halo pressure because p is a pointwise function and the
//      reflection of cell data is a plain copy), soundspeed and viscosity go straight into the dt minimum.
//      17 algorithmic passes -> 9 (reads d0,e0,u0,v0,volume,xarea,yarea; writes p,q; the sound speed is consumed on
//      chip and left unevaluated in memory -- runtime.cu: lazy_soundspeed -- unless the register variant runs).
//      The viscosity halo update is held back and merged into the pressure halo update of pattern P.
//   P  PdV predictor -> ideal_gas(d1,e1) -> [exchange][update_halo]{p} -> revert
//      one kernel: the predicted density/energy live in registers only (revert would overwrite them);
//      19 passes -> 10 or 11 (soundspeed is stored only if something reads it before it is next overwritten).
//   C  accelerate -> PdV corrector -> flux_calc
//      one kernel: a CTA computes the new velocities of its vertex tile (+1 row, +1 column) into shared
//      memory, then the cells and faces of the tile take them from there; 31 passes -> 15.
//   M  advec_mom(xvel1) -> advec_mom(yvel1), same sweep: one launch for both components.
//
// Every fused kernel executes the same per-cell arithmetic, in the same order, as the single-call kernels
// (lagrange.cuh), so every array the host can observe afterwards is bit-identical (tests/test_gpu_run.py,
// tests/test_gpu_fusion.py run both ways and compare all 15 fields including halos).
#include <cstring>

#include "clover_b200.h"
#include "common.cuh"
#include "lagrange.cuh"
#include "tma.cuh"

namespace clv {

// from lagrange.cu
void set_dt_result_seq(double s);
// from runtime.cu
bool tma_enabled();
int sm_count();
bool chunk_registered();
const int* chunk_neighbours();

// ================================================================================================
// T: ideal_gas + viscosity + calc_dt.  Persistent CTAs over 32x8 tiles (as calc_dt), one cell per thread.
// The thread of cell (j,k) evaluates the equation of state for its own cell (pressure, soundspeed) and the
// pressures of its four face neighbours; threads on the rim of the chunk also store the neighbour's
// pressure into the depth-1 halo ring (what the halo update of `pressure` would have put there).
template <bool SAFE>
__device__ __forceinline__ double timestep_cell(double rho, double en, ViscIn& V, DtIn& D, const DtParams& P,
                                                double& p, double& ss, double& q, bool& bad) {
  ideal_gas_cell<SAFE>(rho, en, p, ss, bad);
  q = viscosity_cell<SAFE>(V, bad);
  D.ssp = ss;
  D.visc = q;
  return calc_dt_cell<SAFE>(D, P, bad);
}

template <bool WRITE_SS>
__global__ void __launch_bounds__(BX* BY)
    timestep_kernel(Range r, int pitch, DtParams P, const double* __restrict__ xarea,
                    const double* __restrict__ yarea, const double* __restrict__ celldx,
                    const double* __restrict__ celldy, const double* __restrict__ volume,
                    const double* __restrict__ density0, const double* __restrict__ energy0,
                    double* __restrict__ pressure, double* __restrict__ viscosity,
                    double* __restrict__ soundspeed, const double* __restrict__ xvel0,
                    const double* __restrict__ yvel0, double* __restrict__ partials, unsigned int* ticket,
                    double* __restrict__ out, ReduceTail RT) {
  double m[1] = {P.g_big};
  CLV_PTILES_BEGIN(r, 1)
    const size_t c = idx2(pitch, j, k);
    // all loads first (one batch in flight), then the arithmetic
    const double rho = density0[c], en = energy0[c];
    const double rl = density0[c - 1], rr = density0[c + 1], rb = density0[c - pitch], rt = density0[c + pitch];
    const double el = energy0[c - 1], er = energy0[c + 1], eb = energy0[c - pitch], et = energy0[c + pitch];
    const double u00 = xvel0[c], u10 = xvel0[c + 1], u01 = xvel0[c + pitch], u11 = xvel0[c + pitch + 1];
    const double v00 = yvel0[c], v10 = yvel0[c + 1], v01 = yvel0[c + pitch], v11 = yvel0[c + pitch + 1];
    const double dsx = celldx[j + 1], dsy = celldy[k + 1], dsx1 = celldx[j + 2], dsy1 = celldy[k + 2];
    const double vol = volume[c];
    const double xa0 = xarea[c], xa1 = xarea[c + 1], ya0 = yarea[c], ya1 = yarea[c + pitch];
    if (k + PF_ROWS <= r.k1) {
      const size_t pf = c + (size_t)PF_ROWS * pitch;
      prefetch_l2(density0 + pf); prefetch_l2(energy0 + pf); prefetch_l2(xvel0 + pf); prefetch_l2(yvel0 + pf);
      prefetch_l2(volume + pf); prefetch_l2(xarea + pf); prefetch_l2(yarea + pf);
    }
    // ideal_gas_kernel_c.c:52 for the four neighbours
    const double pl = (1.4 - 1.0) * rl * el, pr = (1.4 - 1.0) * rr * er;
    const double pb = (1.4 - 1.0) * rb * eb, pt = (1.4 - 1.0) * rt * et;
    ViscIn V{u00, u10, u01, u11, v00, v10, v01, v11, dsx, dsy, dsx1, dsy1, pl, pr, pb, pt, rho};
    DtIn D{dsx, dsy, vol, 0.0, 0.0, rho, u00, u10, u01, u11, v00, v10, v01, v11, xa0, xa1, ya0, ya1};
    bool bad = false;
    double p, ss, q;
    double cell_dt = timestep_cell<false>(rho, en, V, D, P, p, ss, q, bad);
    if (bad) cell_dt = timestep_cell<true>(rho, en, V, D, P, p, ss, q, bad);
    if (active) {
      pressure[c] = p;
      viscosity[c] = q;
      if (WRITE_SS) soundspeed[c] = ss;
      if (cell_dt < m[0]) m[0] = cell_dt;
      // depth-1 halo ring of pressure (corners by the corner cells)
      const bool L = (j == r.j0), R_ = (j == r.j1), B = (k == r.k0), T = (k == r.k1);
      if (L) pressure[c - 1] = pl;
      if (R_) pressure[c + 1] = pr;
      if (B) pressure[c - pitch] = pb;
      if (T) pressure[c + pitch] = pt;
      if ((L || R_) && (B || T)) {
        const size_t cc = c + (L ? -1 : 1) + (B ? -(ptrdiff_t)pitch : (ptrdiff_t)pitch);
        pressure[cc] = (1.4 - 1.0) * density0[cc] * energy0[cc];
        // a one-cell-wide or one-cell-high chunk: the same cell is on both rims
        if (L && R_) { const size_t c2 = c + 1 + (B ? -(ptrdiff_t)pitch : (ptrdiff_t)pitch); pressure[c2] = (1.4 - 1.0) * density0[c2] * energy0[c2]; }
        if (B && T) { const size_t c2 = c + (L ? -1 : 1) + (ptrdiff_t)pitch; pressure[c2] = (1.4 - 1.0) * density0[c2] * energy0[c2]; }
        if (L && R_ && B && T) { const size_t c2 = c + 1 + (ptrdiff_t)pitch; pressure[c2] = (1.4 - 1.0) * density0[c2] * energy0[c2]; }
      }
    }
  CLV_PTILES_END
  block_reduce_publish<1, true>(m, partials, ticket, out, P.g_big, RT);
}

// One cell of T for the TMA kernel.  What only a compressing cell needs -- the four neighbour pressures and the
// limiter of viscosity_kernel_c.c:60-104 -- is fetched from the staged tile and evaluated inside that minority branch;
// the dt minimum takes the division-saving form (lagrange.cuh: calc_dt_cell_lean).  Same values as timestep_cell.
struct TsCell {
  double rho, en, u00, u10, u01, u11, v00, v10, v01, v11, vol, xa0, xa1, ya0, ya1, dsx, dsy, dsx1, dsy1;
};
template <bool SAFE>
__device__ __forceinline__ double timestep_cell_lean(const TsCell& C, const double* __restrict__ sd,
                                                     const double* __restrict__ se, int b, int bw, const DtParams& P,
                                                     unsigned mask, double& p, double& ss, double& q, bool& bad) {
  ideal_gas_cell<SAFE>(C.rho, C.en, p, ss, bad);
  const double ugrad = (C.u10 + C.u11) - (C.u00 + C.u01);
  const double vgrad = (C.v01 + C.v11) - (C.v00 + C.v10);
  const double div = C.dsx * ugrad + C.dsy * vgrad;
  q = 0.0;
  if (!(div >= 0.0)) {  // viscosity_kernel_c.c:88: 0 unless compressing
    const double pl = (1.4 - 1.0) * sd[b - 1] * se[b - 1], pr = (1.4 - 1.0) * sd[b + 1] * se[b + 1];
    const double pb = (1.4 - 1.0) * sd[b - bw] * se[b - bw], pt = (1.4 - 1.0) * sd[b + bw] * se[b + bw];
    const ViscIn V{C.u00, C.u10, C.u01, C.u11, C.v00, C.v10, C.v01, C.v11, C.dsx, C.dsy, C.dsx1, C.dsy1, pl, pr, pb, pt, C.rho};
    q = viscosity_cell<SAFE>(V, bad);
  }
  const DtIn D{C.dsx, C.dsy, C.vol, ss, q, C.rho, C.u00, C.u10, C.u01, C.u11, C.v00, C.v10, C.v01, C.v11, C.xa0, C.xa1, C.ya0, C.ya1};
  return calc_dt_cell_lean<SAFE>(D, P, bad, mask);
}

// ---- T with TMA tile staging (tma.cuh): persistent CTAs of 32x8 threads, TT_RPT cells per thread, the seven input
// fields of a 32 x 8*TT_RPT tile arrive as 36 x (8*TT_RPT+2) boxes with corner (j0-2, k0-1).  Same arithmetic as
// timestep_kernel.
// Rows per thread (rows ly, ly+BY, ...; unrolled, so that the division / square-root chains of independent cells
// interleave -- the kernel's top stall is the fixed-latency dependency wait).  Measured on B200 at 3840^2, 2 CTAs/SM
// (profiles/r02_experiment_timestep_rows.txt): 1 row x 4 stages 0.243 ms, 2 rows x 3 stages 0.212, 2 x 2 0.213,
// 2 rows one after the other (unroll 1) 0.232.
#ifndef TT_RPT
#define TT_RPT 2
#endif
constexpr int TT_W = BX, TT_H = BY * TT_RPT, TT_BW = TT_W + 4, TT_BH = TT_H + 2, TT_NARR = 7;
// With one row per thread: (stages, CTAs/SM) = (4,2) 0.243 ms at 127 registers; three CTAs per SM need 80 registers,
// which spills ~140 bytes: (3,3) 0.278, (2,3) 0.256 (profiles/r02_experiment_occupancy.txt)
#ifndef TT_STAGES
#define TT_STAGES 3
#endif
#ifndef TT_CPS
#define TT_CPS 2
#endif
enum { TA_D0 = 0, TA_E0, TA_U0, TA_V0, TA_VOL, TA_XA, TA_YA };
using TimestepRing = TileRing<TT_NARR, TT_BW, TT_BH, TT_STAGES>;
constexpr int TT_SMEM = TimestepRing::BYTES + 128;
static_assert(fits_sm(TT_SMEM, TT_CPS), "timestep: TT_CPS CTAs do not fit one SM");
struct TimestepMaps {
  CUtensorMap m[TT_NARR];
};

template <bool WRITE_SS>
__global__ void __launch_bounds__(BX* BY, TT_CPS)
    timestep_tma_kernel(const __grid_constant__ TimestepMaps M, DtParams P, const double* __restrict__ celldx,
                        const double* __restrict__ celldy, const double* __restrict__ density0,
                        const double* __restrict__ energy0, double* __restrict__ pressure,
                        double* __restrict__ viscosity, double* __restrict__ soundspeed, double* __restrict__ partials,
                        unsigned int* ticket, double* __restrict__ out, int nx, int ny, int pitch, int ntx, int ntiles,
                        const int2* __restrict__ order, Tickets tickets, int dep_start, unsigned long long* trace, ReduceTail RT) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = align128(smem_raw);
  TimestepRing ring;
  ring.init(smem);
  pdl_trigger();
  PdlGate gate(dep_start, trace);
  const int lx = threadIdx.x, ly = threadIdx.y;
  const bool leader = (lx == 0 && ly == 0);
  const int G = gridDim.x;
  double m[1] = {P.g_big};
  auto issue_tile = [&](int stage, int2 xy) {
    const int j0 = 1 + xy.x * TT_W, k0 = 1 + xy.y * TT_H;
    ring.issue(M.m, stage, j0 - 2 + XOFF, k0 - 1 + 1);
  };
  __shared__ int s_tile[TT_STAGES];
  __shared__ int2 s_xy[TT_STAGES];
  __shared__ int s_q[8];
  TileQueue<TT_STAGES> queue(tickets, ntiles, order, s_tile, s_xy, s_q);
  const bool sched = (lx == 0 && ly == 1);  // lane 0 of warp 1 drives the tile queue (tma.cuh)
  if (sched) queue.prime_all();
  __syncthreads();
  if (leader) {
#pragma unroll
    for (int s = 0; s < TT_STAGES - 1; ++s) {
      if (s_tile[s] < ntiles) {
        gate.need(s_tile[s]);
        issue_tile(s, s_xy[s]);
      }
    }
  }
  // this iteration's tile sits in registers: it was read from the ring's table during the previous iteration (slot
  // stage+1 is refilled STAGES-1 >= 1 iterations before it is read, i.e. behind at least one barrier)
  int t = s_tile[0];
  int2 cur = s_xy[0];
  __syncthreads();  // slot 0 has been read: the scheduler may refill it
  for (int it = 0; t < ntiles; ++it) {
    const int stage = it % TT_STAGES;
    if (sched) queue.step(stage);
    gate.need(t);  // rim tiles wait for the halo kernel (their loads, issued STAGES-1 tiles ahead, waited in the leader)
    if (leader) {
      const int ns = (stage + TT_STAGES - 1) % TT_STAGES;
      const int tn = s_tile[ns];
      if (tn < ntiles) {
        gate.need(tn);
        issue_tile(ns, s_xy[ns]);
      }
    }
    const int j = 1 + cur.x * TT_W + lx, k_top = 1 + cur.y * TT_H + ly;
    const int jc = j <= nx ? j : nx;  // 1-D geometry of the threads beyond the chunk
    const double dsx = celldx[jc + 1], dsx1 = celldx[jc + 2];
    double dsy_r[TT_RPT], dsy1_r[TT_RPT];
#pragma unroll
    for (int r = 0; r < TT_RPT; ++r) {
      const int kr = k_top + r * BY, kc = kr <= ny ? kr : ny;
      dsy_r[r] = celldy[kc + 1];
      dsy1_r[r] = celldy[kc + 2];
    }
    ring.wait(stage, (uint32_t)((it / TT_STAGES) & 1));
    const int t_nx = s_tile[(stage + 1) % TT_STAGES];
    const int2 xy_nx = s_xy[(stage + 1) % TT_STAGES];
    const double* __restrict__ sd = ring.tile(stage, TA_D0);
    const double* __restrict__ se = ring.tile(stage, TA_E0);
    const double* __restrict__ su = ring.tile(stage, TA_U0);
    const double* __restrict__ sv = ring.tile(stage, TA_V0);
    const double* __restrict__ svol = ring.tile(stage, TA_VOL);
    const double* __restrict__ sxa = ring.tile(stage, TA_XA);
    const double* __restrict__ sya = ring.tile(stage, TA_YA);
#pragma unroll
    for (int r = 0; r < TT_RPT; ++r) {
    const int lyr = ly + r * BY, k = k_top + r * BY;
    const bool active = j <= nx && k <= ny;
    const double dsy = dsy_r[r], dsy1 = dsy1_r[r];
    const int b = (lyr + 1) * TT_BW + lx + 2;
    TsCell C;
    C.rho = sd[b]; C.en = se[b];
    C.u00 = su[b]; C.u10 = su[b + 1]; C.u01 = su[b + TT_BW]; C.u11 = su[b + TT_BW + 1];
    C.v00 = sv[b]; C.v10 = sv[b + 1]; C.v01 = sv[b + TT_BW]; C.v11 = sv[b + TT_BW + 1];
    C.vol = svol[b];
    C.xa0 = sxa[b]; C.xa1 = sxa[b + 1]; C.ya0 = sya[b]; C.ya1 = sya[b + TT_BW];
    C.dsx = dsx; C.dsy = dsy; C.dsx1 = dsx1; C.dsy1 = dsy1;
    if (active) {
      bool bad = false;
      double p, ss, q;
      const unsigned mask = __activemask();
      double cell_dt = timestep_cell_lean<false>(C, sd, se, b, TT_BW, P, mask, p, ss, q, bad);
      if (bad) cell_dt = timestep_cell_lean<true>(C, sd, se, b, TT_BW, P, mask, p, ss, q, bad);
      const size_t c = idx2(pitch, j, k);
      pressure[c] = p;
      viscosity[c] = q;
      if (WRITE_SS) soundspeed[c] = ss;
      if (cell_dt < m[0]) m[0] = cell_dt;
      // depth-1 halo ring of pressure (corners by the corner cells): ideal_gas_kernel_c.c:52 for the neighbour cell
      const bool L = (j == 1), R_ = (j == nx), B = (k == 1), T = (k == ny);
      if (L | R_ | B | T) {
        if (L) pressure[c - 1] = (1.4 - 1.0) * sd[b - 1] * se[b - 1];
        if (R_) pressure[c + 1] = (1.4 - 1.0) * sd[b + 1] * se[b + 1];
        if (B) pressure[c - pitch] = (1.4 - 1.0) * sd[b - TT_BW] * se[b - TT_BW];
        if (T) pressure[c + pitch] = (1.4 - 1.0) * sd[b + TT_BW] * se[b + TT_BW];
        if ((L || R_) && (B || T)) {
          const size_t cc = c + (L ? -1 : 1) + (B ? -(ptrdiff_t)pitch : (ptrdiff_t)pitch);
          pressure[cc] = (1.4 - 1.0) * density0[cc] * energy0[cc];
          // a one-cell-wide or one-cell-high chunk: the same cell is on both rims
          if (L && R_) { const size_t c2 = c + 1 + (B ? -(ptrdiff_t)pitch : (ptrdiff_t)pitch); pressure[c2] = (1.4 - 1.0) * density0[c2] * energy0[c2]; }
          if (B && T) { const size_t c2 = c + (L ? -1 : 1) + (ptrdiff_t)pitch; pressure[c2] = (1.4 - 1.0) * density0[c2] * energy0[c2]; }
          if (L && R_ && B && T) { const size_t c2 = c + 1 + (ptrdiff_t)pitch; pressure[c2] = (1.4 - 1.0) * density0[c2] * energy0[c2]; }
        }
      }
    }
    }  // rows of this thread
    __syncthreads();  // the stage (read on demand above: neighbour pressures of compressing / rim cells) can be refilled
    t = t_nx;
    cur = xy_nx;
  }
  gate.finish();
  if (sched) queue.leave();
  block_reduce_publish<1, true>(m, partials, ticket, out, P.g_big, RT);
}

// ================================================================================================
// P: PdV predictor (PdV_kernel_c.c:63-113) + ideal_gas on the predicted state (ideal_gas_kernel_c.c:48-59).
// The predicted density1/energy1 are not stored: revert (revert_kernel_c.c:46-62) replaces them right after.
constexpr int NR_PRED = 2;
template <bool WRITE_SS>
__global__ void __launch_bounds__(BX* BY)
    pdv_predict_eos_kernel(Range r, int pitch, double dt, const double* __restrict__ xarea,
                           const double* __restrict__ yarea, const double* __restrict__ volume,
                           const double* __restrict__ density0, const double* __restrict__ energy0,
                           double* __restrict__ pressure, const double* __restrict__ viscosity,
                           double* __restrict__ soundspeed, const double* __restrict__ xvel0,
                           const double* __restrict__ yvel0) {
  CLV_ROWS_BEGIN(r, NR_PRED)
    const size_t c = idx2(pitch, j, k);
    const double x00 = xvel0[c], x10 = xvel0[c + 1], x01 = xvel0[c + pitch], x11 = xvel0[c + pitch + 1];
    const double y00 = yvel0[c], y10 = yvel0[c + 1], y01 = yvel0[c + pitch], y11 = yvel0[c + pitch + 1];
    const double vol = volume[c], rho0 = density0[c], pres = pressure[c], visc = viscosity[c], en0 = energy0[c];
    const double xa0 = xarea[c], xa1 = xarea[c + 1], ya0 = yarea[c], ya1 = yarea[c + pitch];
    if (k + PF_ROWS <= r.k1) {
      const size_t pf = c + (size_t)PF_ROWS * pitch;
      prefetch_l2(xvel0 + pf); prefetch_l2(yvel0 + pf); prefetch_l2(volume + pf); prefetch_l2(density0 + pf);
      prefetch_l2(pressure + pf); prefetch_l2(viscosity + pf); prefetch_l2(energy0 + pf); prefetch_l2(xarea + pf);
      prefetch_l2(yarea + pf);
    }
    const double left = xa0 * (x00 + x01 + x00 + x01) * 0.25 * dt * 0.5;
    const double right = xa1 * (x10 + x11 + x10 + x11) * 0.25 * dt * 0.5;
    const double bottom = ya0 * (y00 + y10 + y00 + y10) * 0.25 * dt * 0.5;
    const double top = ya1 * (y01 + y11 + y01 + y11) * 0.25 * dt * 0.5;
    const double total = right - left + top - bottom;
    const double vc = vol / (vol + total);
    const double recip = 1.0 / vol;
    const double de = (pres / rho0 + ddiv(visc, rho0)) * total * recip;
    const double e1 = en0 - de;
    const double d1 = rho0 * vc;
    bool bad = false;
    double p, ss;
    ideal_gas_cell<false>(d1, e1, p, ss, bad);
    if (bad) ideal_gas_cell<true>(d1, e1, p, ss, bad);
    if (active) {
      pressure[c] = p;
      if (WRITE_SS) soundspeed[c] = ss;
    }
  CLV_ROWS_END
}

// ---- P with TMA tile staging: 32x8 threads, PT_RPT (one) cell per thread, nine 34 x (8*PT_RPT+1) boxes with corner
// (j0, k0); two cells per thread make the kernel 4 % faster and the step 1 % slower (r02_experiment_timestep_rows.txt)
#ifndef PT_RPT
#define PT_RPT 1  // tile rows per thread (rows ly, ly+BY, ...), unrolled
#endif
constexpr int PT_W = BX, PT_H = BY * PT_RPT, PT_BW = PT_W + 2, PT_BH = PT_H + 1, PT_NARR = 9;
// measured on B200 at 3840^2: (stages, CTAs/SM) = (4,2) 0.202 ms, (2,4) 0.217, (3,3) 0.228
#ifndef PT_STAGES
#define PT_STAGES 4
#endif
#ifndef PT_CPS
#define PT_CPS 2
#endif
enum { PA_XAREA = 0, PA_YAREA, PA_VOLUME, PA_DENSITY0, PA_ENERGY0, PA_PRESSURE, PA_VISCOSITY, PA_XVEL0, PA_YVEL0 };
using PredictRing = TileRing<PT_NARR, PT_BW, PT_BH, PT_STAGES>;
constexpr int PT_SMEM = PredictRing::BYTES + 128;
static_assert(fits_sm(PT_SMEM, PT_CPS), "pdv_predict: PT_CPS CTAs do not fit one SM");
struct PredictMaps {
  CUtensorMap m[PT_NARR];
};

template <bool WRITE_SS>
__global__ void __launch_bounds__(BX* BY, PT_CPS)
    pdv_predict_eos_tma_kernel(const __grid_constant__ PredictMaps M, double dt, double* __restrict__ pressure,
                               double* __restrict__ soundspeed, int nx, int ny, int pitch, int ntx, int ntiles,
                        const int2* __restrict__ order, Tickets tickets, int dep_start, unsigned long long* trace) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = align128(smem_raw);
  PredictRing ring;
  ring.init(smem);
  pdl_trigger();
  PdlGate gate(dep_start, trace);
  const int lx = threadIdx.x, ly = threadIdx.y;
  const bool leader = (lx == 0 && ly == 0);
  const int G = gridDim.x;
  auto issue_tile = [&](int stage, int2 xy) {
    const int j0 = 1 + xy.x * PT_W, k0 = 1 + xy.y * PT_H;
    ring.issue(M.m, stage, j0 + XOFF, k0 + 1);
  };
  __shared__ int s_tile[PT_STAGES];
  __shared__ int2 s_xy[PT_STAGES];
  __shared__ int s_q[8];
  TileQueue<PT_STAGES> queue(tickets, ntiles, order, s_tile, s_xy, s_q);
  const bool sched = (lx == 0 && ly == 1);  // lane 0 of warp 1 drives the tile queue (tma.cuh)
  if (sched) queue.prime_all();
  __syncthreads();
  if (leader) {
#pragma unroll
    for (int s = 0; s < PT_STAGES - 1; ++s) {
      if (s_tile[s] < ntiles) {
        gate.need(s_tile[s]);
        issue_tile(s, s_xy[s]);
      }
    }
  }
  // this iteration's tile sits in registers: it was read from the ring's table during the previous iteration (slot
  // stage+1 is refilled STAGES-1 >= 1 iterations before it is read, i.e. behind at least one barrier)
  int t = s_tile[0];
  int2 cur = s_xy[0];
  __syncthreads();  // slot 0 has been read: the scheduler may refill it
  for (int it = 0; t < ntiles; ++it) {
    const int stage = it % PT_STAGES;
    if (sched) queue.step(stage);
    gate.need(t);
    if (leader) {
      const int ns = (stage + PT_STAGES - 1) % PT_STAGES;
      const int tn = s_tile[ns];
      if (tn < ntiles) {
        gate.need(tn);
        issue_tile(ns, s_xy[ns]);
      }
    }
    const int j = 1 + cur.x * PT_W + lx, k_top = 1 + cur.y * PT_H + ly;
    ring.wait(stage, (uint32_t)((it / PT_STAGES) & 1));
    const int t_nx = s_tile[(stage + 1) % PT_STAGES];
    const int2 xy_nx = s_xy[(stage + 1) % PT_STAGES];
    const double* __restrict__ sxa = ring.tile(stage, PA_XAREA);
    const double* __restrict__ sya = ring.tile(stage, PA_YAREA);
    const double* __restrict__ sx = ring.tile(stage, PA_XVEL0);
    const double* __restrict__ sy = ring.tile(stage, PA_YVEL0);
    double x00[PT_RPT], x10[PT_RPT], x01[PT_RPT], x11[PT_RPT], y00[PT_RPT], y10[PT_RPT], y01[PT_RPT], y11[PT_RPT];
    double vol[PT_RPT], rho0[PT_RPT], pres[PT_RPT], visc[PT_RPT], en0[PT_RPT], xa0[PT_RPT], xa1[PT_RPT], ya0[PT_RPT], ya1[PT_RPT];
#pragma unroll
    for (int r = 0; r < PT_RPT; ++r) {
      const int b = (ly + r * BY) * PT_BW + lx;
      x00[r] = sx[b]; x10[r] = sx[b + 1]; x01[r] = sx[b + PT_BW]; x11[r] = sx[b + PT_BW + 1];
      y00[r] = sy[b]; y10[r] = sy[b + 1]; y01[r] = sy[b + PT_BW]; y11[r] = sy[b + PT_BW + 1];
      vol[r] = ring.tile(stage, PA_VOLUME)[b]; rho0[r] = ring.tile(stage, PA_DENSITY0)[b];
      pres[r] = ring.tile(stage, PA_PRESSURE)[b]; visc[r] = ring.tile(stage, PA_VISCOSITY)[b];
      en0[r] = ring.tile(stage, PA_ENERGY0)[b];
      xa0[r] = sxa[b]; xa1[r] = sxa[b + 1]; ya0[r] = sya[b]; ya1[r] = sya[b + PT_BW];
    }
    __syncthreads();  // everything this tile needs is in registers: the stage can be refilled
#pragma unroll
    for (int r = 0; r < PT_RPT; ++r) {
      const int k = k_top + r * BY;
      if (j <= nx && k <= ny) {
        // PdV_kernel_c.c:63-113 (predictor), ideal_gas_kernel_c.c:48-59 on the predicted state
        const double left = xa0[r] * (x00[r] + x01[r] + x00[r] + x01[r]) * 0.25 * dt * 0.5;
        const double right = xa1[r] * (x10[r] + x11[r] + x10[r] + x11[r]) * 0.25 * dt * 0.5;
        const double bottom = ya0[r] * (y00[r] + y10[r] + y00[r] + y10[r]) * 0.25 * dt * 0.5;
        const double top = ya1[r] * (y01[r] + y11[r] + y01[r] + y11[r]) * 0.25 * dt * 0.5;
        const double total = right - left + top - bottom;
        const double vc = vol[r] / (vol[r] + total);
        const double recip = 1.0 / vol[r];
        const double de = (pres[r] / rho0[r] + ddiv(visc[r], rho0[r])) * total * recip;
        const double e1 = en0[r] - de;
        const double d1 = rho0[r] * vc;
        bool bad = false;
        double p, ss;
        ideal_gas_cell<false>(d1, e1, p, ss, bad);
        if (bad) ideal_gas_cell<true>(d1, e1, p, ss, bad);
        const size_t c = idx2(pitch, j, k);
        pressure[c] = p;
        if (WRITE_SS) soundspeed[c] = ss;
      }
    }
    t = t_nx;
    cur = xy_nx;
  }
  gate.finish();
  if (sched) queue.leave();
}

// ================================================================================================
// C: accelerate (accelerate_kernel_c.c:56-95) + PdV corrector (PdV_kernel_c.c:115-167) + flux_calc
// (flux_calc_kernel_c.c:49-73).  A CTA owns the slots (j,k) of a CT_W x CT_H tile of 1..nx+1 x 1..ny+1;
// slot (j,k) = vertex (j,k), x-face (j,k), y-face (j,k) and cell (j,k).  Phase 1: new velocities of the
// tile's vertices plus one extra row and column (the upper/right vertices of the last cells) go to shared
// memory together with the old ones; phase 2: faces and cells read their two / four vertices from there.
constexpr int CT_W = 64, CT_H = 16, CT_BY = 4, CT_NR = CT_H / CT_BY;
struct CorrectArgs {
  const double *xarea, *yarea, *volume, *density0, *energy0, *pressure, *viscosity, *xvel0, *yvel0;
  double *xvel1, *yvel1, *density1, *energy1, *vol_flux_x, *vol_flux_y;
};
__device__ __forceinline__ void accel_vertex(const CorrectArgs& A, size_t c11, int pitch, double dt, double& xv0,
                                             double& yv0, double& xv, double& yv) {
  const size_t c01 = c11 - 1, c10 = c11 - pitch, c00 = c10 - 1;
  const double d00 = A.density0[c00], d10 = A.density0[c10], d11 = A.density0[c11], d01 = A.density0[c01];
  const double w00 = A.volume[c00], w10 = A.volume[c10], w11 = A.volume[c11], w01 = A.volume[c01];
  const double xa1 = A.xarea[c11], xa0 = A.xarea[c10];
  const double ya1 = A.yarea[c11], ya0 = A.yarea[c01];
  const double p11 = A.pressure[c11], p01 = A.pressure[c01], p10 = A.pressure[c10], p00 = A.pressure[c00];
  const double q11 = A.viscosity[c11], q01 = A.viscosity[c01], q10 = A.viscosity[c10], q00 = A.viscosity[c00];
  xv0 = A.xvel0[c11];
  yv0 = A.yvel0[c11];
  const double nodal_mass = (d00 * w00 + d10 * w10 + d11 * w11 + d01 * w01) * 0.25;
  const double s = 0.5 * dt / nodal_mass;
  xv = xv0 - s * (xa1 * (p11 - p01) + xa0 * (p10 - p00));
  yv = yv0 - s * (ya1 * (p11 - p10) + ya0 * (p01 - p00));
  xv = xv - s * (xa1 * (q11 - q01) + xa0 * (q10 - q00));
  yv = yv - s * (ya1 * (q11 - q10) + ya0 * (q01 - q00));
}

__global__ void __launch_bounds__(CT_W* CT_BY)
    lagrange_correct_kernel(CorrectArgs A, int nx, int ny, int pitch, double dt) {
  __shared__ double su0[CT_H + 1][CT_W + 1], sv0[CT_H + 1][CT_W + 1], su1[CT_H + 1][CT_W + 1], sv1[CT_H + 1][CT_W + 1];
  const int lx = threadIdx.x, ty = threadIdx.y;
  const int j0 = 1 + (int)blockIdx.x * CT_W, k0 = 1 + (int)blockIdx.y * CT_H;
  const int jmax = nx + 1, kmax = ny + 1;
  // ---- phase 1: vertices --------------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < CT_NR; ++i) {
    const int ly = ty + i * CT_BY;
    const int jr = j0 + lx, kr = k0 + ly;
    const int j = jr <= jmax ? jr : jmax, k = kr <= kmax ? kr : kmax;  // clamped: always safe to load
    const size_t c = idx2(pitch, j, k);
    if (kr + PF_ROWS <= kmax) {
      const size_t pf = c + (size_t)PF_ROWS * pitch;
      prefetch_l2(A.density0 + pf); prefetch_l2(A.volume + pf); prefetch_l2(A.pressure + pf);
      prefetch_l2(A.viscosity + pf); prefetch_l2(A.xarea + pf); prefetch_l2(A.yarea + pf);
      prefetch_l2(A.xvel0 + pf); prefetch_l2(A.yvel0 + pf); prefetch_l2(A.energy0 + pf);
    }
    double xv0, yv0, xv, yv;
    accel_vertex(A, c, pitch, dt, xv0, yv0, xv, yv);
    su0[ly][lx] = xv0; sv0[ly][lx] = yv0; su1[ly][lx] = xv; sv1[ly][lx] = yv;
    if (jr <= jmax && kr <= kmax) {
      A.xvel1[c] = xv;
      A.yvel1[c] = yv;
    }
  }
  {
    // the extra row (ly = CT_H, lx = 0..CT_W) and column (lx = CT_W, ly = 0..CT_H-1): owned by the neighbour tiles
    const int t = ty * CT_W + lx;
    if (t < CT_W + 1 + CT_H) {
      const int ex = t <= CT_W ? t : CT_W, ey = t <= CT_W ? CT_H : t - (CT_W + 1);
      const int jr = j0 + ex, kr = k0 + ey;
      const int j = jr <= jmax ? jr : jmax, k = kr <= kmax ? kr : kmax;
      double xv0, yv0, xv, yv;
      accel_vertex(A, idx2(pitch, j, k), pitch, dt, xv0, yv0, xv, yv);
      su0[ey][ex] = xv0; sv0[ey][ex] = yv0; su1[ey][ex] = xv; sv1[ey][ex] = yv;
    }
  }
  __syncthreads();
  // ---- phase 2: faces and cells ----------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < CT_NR; ++i) {
    const int ly = ty + i * CT_BY;
    const int j = j0 + lx, k = k0 + ly;
    if (j > jmax || k > kmax) continue;
    const size_t c = idx2(pitch, j, k);
    const double x00 = su0[ly][lx], x10 = su0[ly][lx + 1], x01 = su0[ly + 1][lx], x11 = su0[ly + 1][lx + 1];
    const double y00 = sv0[ly][lx], y10 = sv0[ly][lx + 1], y01 = sv0[ly + 1][lx], y11 = sv0[ly + 1][lx + 1];
    const double a00 = su1[ly][lx], a10 = su1[ly][lx + 1], a01 = su1[ly + 1][lx], a11 = su1[ly + 1][lx + 1];
    const double b00 = sv1[ly][lx], b10 = sv1[ly][lx + 1], b01 = sv1[ly + 1][lx], b11 = sv1[ly + 1][lx + 1];
    const double xa0 = A.xarea[c], ya0 = A.yarea[c];
    // flux_calc_kernel_c.c:55-58, :67-70
    if (k <= ny) A.vol_flux_x[c] = 0.25 * dt * xa0 * (x00 + x01 + a00 + a01);
    if (j <= nx) A.vol_flux_y[c] = 0.25 * dt * ya0 * (y00 + y10 + b00 + b10);
    if (j <= nx && k <= ny) {
      // PdV_kernel_c.c:117-163
      const double xa1 = A.xarea[c + 1], ya1 = A.yarea[c + pitch];
      const double vol = A.volume[c], rho0 = A.density0[c], pres = A.pressure[c], visc = A.viscosity[c];
      const double en0 = A.energy0[c];
      const double left = xa0 * (x00 + x01 + a00 + a01) * 0.25 * dt;
      const double right = xa1 * (x10 + x11 + a10 + a11) * 0.25 * dt;
      const double bottom = ya0 * (y00 + y10 + b00 + b10) * 0.25 * dt;
      const double top = ya1 * (y01 + y11 + b01 + b11) * 0.25 * dt;
      const double total = right - left + top - bottom;
      const double vc = vol / (vol + total);
      const double recip = 1.0 / vol;
      const double de = (pres / rho0 + ddiv(visc, rho0)) * total * recip;
      A.energy1[c] = en0 - de;
      A.density1[c] = rho0 * vc;
    }
  }
}

// ---- C with TMA tile staging (tma.cuh) ------------------------------------------------------------------------
// Persistent CTAs; a tile is LT_W x LT_H slots.  The nine input fields arrive as (LT_W+4) x (LT_H+2) boxes with
// corner (j0-2, k0-1): cells j0-1..j0+LT_W feed the nodal masses of vertices j0..j0+LT_W, and the same box holds
// the vertices / faces of the tile plus the extra row and column.  The arithmetic is that of
// lagrange_correct_kernel, statement for statement.
constexpr int LT_H = 8, LT_BH = LT_H + 2, LT_NARR = 9;
#ifndef LC_W
#define LC_W 64
#endif
#ifndef LC_CPS
#define LC_CPS 2
#endif
#ifndef LC_FASTDIV
#define LC_FASTDIV 1
#endif
constexpr int LT_OX = 2;  // box corner at j0-2: TMA needs a 16-byte aligned start, i.e. an even dim-0 coordinate (tma.cuh)
enum { LA_XAREA = 0, LA_YAREA, LA_VOLUME, LA_DENSITY0, LA_ENERGY0, LA_PRESSURE, LA_VISCOSITY, LA_XVEL0, LA_YVEL0 };
struct CorrectMaps {
  CUtensorMap m[LT_NARR];
};
struct CorrectOut {
  double *xvel1, *yvel1, *density1, *energy1, *vol_flux_x, *vol_flux_y;
};

// RPT = slot rows per thread (independent dependency chains per thread), STAGES = ring depth, CPS = CTAs per SM.
// W = tile width in slots (even), box = (W+4) x (LT_H+2).
template <int W, int RPT, int STAGES, int CPS>
struct CorrectCfg {
  static constexpr int NT = W * LT_H / RPT;
  static constexpr int BW = W + 4, VW = W + 1;
  using Ring = TileRing<LT_NARR, BW, LT_BH, STAGES>;
  static constexpr int NVERT = VW * (LT_H + 1);
  static constexpr int VPT = (NVERT + NT - 1) / NT;  // vertices per thread
  static constexpr int SMEM = Ring::BYTES + 2 * NVERT * 8 + 128;
  static_assert(fits_sm(SMEM, CPS), "lagrange_correct: CPS CTAs of this shape do not fit one SM");
};

template <int W, int RPT, int STAGES, int CPS>
__global__ void __launch_bounds__(W* LT_H / RPT, CPS)
    lagrange_correct_tma_kernel(const __grid_constant__ CorrectMaps M, CorrectOut O, int nx, int ny, int pitch,
                                double dt, int ntx, int ntiles, const int2* __restrict__ order, Tickets tickets, int dep_start,
                                unsigned long long* trace) {
  using Cfg = CorrectCfg<W, RPT, STAGES, CPS>;
  constexpr int NT = Cfg::NT, VPT = Cfg::VPT, NVERT = Cfg::NVERT, ROWS = LT_H / RPT, LT_W = W, LT_BW = Cfg::BW, LT_VW = Cfg::VW;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = align128(smem_raw);
  typename Cfg::Ring ring;
  ring.init(smem);
  pdl_trigger();
  PdlGate gate(dep_start, trace);
  double* __restrict__ su1 = reinterpret_cast<double*>(smem + Cfg::Ring::BYTES);
  double* __restrict__ sv1 = su1 + NVERT;
  const int tid = threadIdx.x, lx = tid % LT_W, ty = tid / LT_W;
  const int G = gridDim.x;
  const int jmax = nx + 1, kmax = ny + 1;
  // box corner of tile t in tensor coordinates: element (j,k) is at (j + XOFF, k + 1)
  auto issue_tile = [&](int stage, int2 xy) {
    const int j0 = 1 + xy.x * LT_W, k0 = 1 + xy.y * LT_H;
    ring.issue(M.m, stage, j0 - LT_OX + XOFF, k0 - 1 + 1);
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
  for (int it = 0;; ++it) {
    const int stage = it % STAGES;
    const int t = s_tile[stage];
    if (t >= ntiles) break;
    const int2 cur = s_xy[stage];
    gate.need(t);
    if (tid == 0) {  // the stage it refills was released by the barrier that ended iteration it-1
      const int ns = (stage + STAGES - 1) % STAGES;
      const int tn = s_tile[ns];
      if (tn < ntiles) {
        gate.need(tn);
        issue_tile(ns, s_xy[ns]);
      }
    }
    const int j0 = 1 + cur.x * LT_W, k0 = 1 + cur.y * LT_H;
    ring.wait(stage, (uint32_t)((it / STAGES) & 1));
    const double* __restrict__ sxa = ring.tile(stage, LA_XAREA);
    const double* __restrict__ sya = ring.tile(stage, LA_YAREA);
    const double* __restrict__ svol = ring.tile(stage, LA_VOLUME);
    const double* __restrict__ sd0 = ring.tile(stage, LA_DENSITY0);
    const double* __restrict__ se0 = ring.tile(stage, LA_ENERGY0);
    const double* __restrict__ sp = ring.tile(stage, LA_PRESSURE);
    const double* __restrict__ sq = ring.tile(stage, LA_VISCOSITY);
    const double* __restrict__ su0 = ring.tile(stage, LA_XVEL0);
    const double* __restrict__ sv0 = ring.tile(stage, LA_YVEL0);
    // ---- phase 1: new velocities of the (LT_W+1) x (LT_H+1) vertices (accelerate_kernel_c.c:56-95) -------------
    {
      double xv[VPT], yv[VPT];
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int v = tid + i * NT;
        const int vc = v < NVERT ? v : NVERT - 1;
        const int vx = vc % LT_VW, vy = vc / LT_VW;
        const int c11 = (vy + 1) * LT_BW + vx + LT_OX, c01 = c11 - 1, c10 = c11 - LT_BW, c00 = c10 - 1;
        const double d00 = sd0[c00], d10 = sd0[c10], d11 = sd0[c11], d01 = sd0[c01];
        const double w00 = svol[c00], w10 = svol[c10], w11 = svol[c11], w01 = svol[c01];
        const double xa1 = sxa[c11], xa0 = sxa[c10];
        const double ya1 = sya[c11], ya0 = sya[c01];
        const double p11 = sp[c11], p01 = sp[c01], p10 = sp[c10], p00 = sp[c00];
        const double q11 = sq[c11], q01 = sq[c01], q10 = sq[c10], q00 = sq[c00];
        const double xv0 = su0[c11], yv0 = sv0[c11];
        const double nodal_mass = (d00 * w00 + d10 * w10 + d11 * w11 + d01 * w01) * 0.25;
        const double s = 0.5 * dt / nodal_mass;  // (the branch-free sequence was measured slower here: 0.322 vs 0.319 ms)
        double x = xv0 - s * (xa1 * (p11 - p01) + xa0 * (p10 - p00));
        double y = yv0 - s * (ya1 * (p11 - p10) + ya0 * (p01 - p00));
        xv[i] = x - s * (xa1 * (q11 - q01) + xa0 * (q10 - q00));
        yv[i] = y - s * (ya1 * (q11 - q10) + ya0 * (q01 - q00));
      }
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int v = tid + i * NT;
        if (v < NVERT) {
          su1[v] = xv[i];
          sv1[v] = yv[i];
          const int vx = v % LT_VW, vy = v / LT_VW;
          const int j = j0 + vx, k = k0 + vy;
          if (vx < LT_W && vy < LT_H && j <= jmax && k <= kmax) {
            const size_t c = idx2(pitch, j, k);
            O.xvel1[c] = xv[i];
            O.yvel1[c] = yv[i];
          }
        }
      }
    }
    __syncthreads();
    if (sched) queue.step(stage);  // (everybody has read this iteration's table slot; the leader reads the new entry after the next barrier)
    // ---- phase 2: faces (flux_calc_kernel_c.c:55-58, :67-70) and cells (PdV_kernel_c.c:117-163) ------------------
    {
      double fx[RPT], fy[RPT], e1[RPT], d1[RPT];
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int ly = ty + r * ROWS;
        const int b = (ly + 1) * LT_BW + lx + LT_OX;  // slot (j,k) inside the box
        const int v = ly * LT_VW + lx;                // vertex (j,k) inside su1/sv1
        const double x00 = su0[b], x10 = su0[b + 1], x01 = su0[b + LT_BW], x11 = su0[b + LT_BW + 1];
        const double y00 = sv0[b], y10 = sv0[b + 1], y01 = sv0[b + LT_BW], y11 = sv0[b + LT_BW + 1];
        const double a00 = su1[v], a10 = su1[v + 1], a01 = su1[v + LT_VW], a11 = su1[v + LT_VW + 1];
        const double b00 = sv1[v], b10 = sv1[v + 1], b01 = sv1[v + LT_VW], b11 = sv1[v + LT_VW + 1];
        const double xa0 = sxa[b], ya0 = sya[b];
        const double xa1 = sxa[b + 1], ya1 = sya[b + LT_BW];
        const double vol = svol[b], rho0 = sd0[b], pres = sp[b], visc = sq[b], en0 = se0[b];
        fx[r] = 0.25 * dt * xa0 * (x00 + x01 + a00 + a01);
        fy[r] = 0.25 * dt * ya0 * (y00 + y10 + b00 + b10);
        const double left = xa0 * (x00 + x01 + a00 + a01) * 0.25 * dt;
        const double right = xa1 * (x10 + x11 + a10 + a11) * 0.25 * dt;
        const double bottom = ya0 * (y00 + y10 + b00 + b10) * 0.25 * dt;
        const double top = ya1 * (y01 + y11 + b01 + b11) * 0.25 * dt;
        const double total = right - left + top - bottom;
#if LC_FASTDIV
        // the three independent quotients of a cell through the branch-free IEEE sequence (common.cuh: Math<false>):
        // with the operators each division ends a basic block and the three run strictly one after the other
        bool bad = false;
        double vc = Math<false>::div(vol, vol + total, bad);
        double recip = Math<false>::rcp(vol, bad);
        double pr = Math<false>::div(pres, rho0, bad);
        if (bad && j0 + lx <= nx && k0 + ly <= ny) {  // (slots beyond the chunk hold zeros: never stored, never re-run)
          vc = vol / (vol + total);
          recip = 1.0 / vol;
          pr = pres / rho0;
        }
        const double de = (pr + ddiv(visc, rho0)) * total * recip;
#else
        const double vc = vol / (vol + total);
        const double recip = 1.0 / vol;
        const double de = (pres / rho0 + ddiv(visc, rho0)) * total * recip;
#endif
        e1[r] = en0 - de;
        d1[r] = rho0 * vc;
      }
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int j = j0 + lx, k = k0 + ty + r * ROWS;
        if (j <= jmax && k <= kmax) {
          const size_t c = idx2(pitch, j, k);
          if (k <= ny) O.vol_flux_x[c] = fx[r];
          if (j <= nx) O.vol_flux_y[c] = fy[r];
          if (j <= nx && k <= ny) {
            O.energy1[c] = e1[r];
            O.density1[c] = d1[r];
          }
        }
      }
    }
    __syncthreads();  // stage and su1/sv1 are free again
  }
  gate.finish();
  if (sched) queue.leave();
}

template <int W, int RPT, int STAGES, int CPS>
static void launch_correct_tma(const CorrectArgs& A, const Grid& g, double dt) {
  using Cfg = CorrectCfg<W, RPT, STAGES, CPS>;
  constexpr int LT_W = W;
  CorrectMaps M;
  const double* in[LT_NARR] = {A.xarea, A.yarea, A.volume, A.density0, A.energy0, A.pressure, A.viscosity, A.xvel0, A.yvel0};
  for (int a = 0; a < LT_NARR; ++a) M.m[a] = *tensor_map_for(g, in[a], Cfg::BW, LT_BH);
  const CorrectOut O{A.xvel1, A.yvel1, A.density1, A.energy1, A.vol_flux_x, A.vol_flux_y};
  static bool configured = false;
  if (!configured) {
    CLV_CUDA(cudaFuncSetAttribute(lagrange_correct_tma_kernel<W, RPT, STAGES, CPS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  const int ntx = (g.nx + 1 + LT_W - 1) / LT_W, nty = (g.ny + 1 + LT_H - 1) / LT_H;
  const int ntiles = ntx * nty;
  const int cap = sm_count() * CPS;
  const int ctas = ntiles < cap ? ntiles : cap;
  const TileOrder ord = tile_order_split(ntx, nty, LT_W, LT_H, LT_OX, Cfg::BW - LT_OX - LT_W, 1, LT_BH - 1 - LT_H, g.nx, g.ny);
  launch_pdl(lagrange_correct_tma_kernel<W, RPT, STAGES, CPS>, dim3(ctas), dim3(Cfg::NT), Cfg::SMEM, stream(), M, O, g.nx, g.ny,
             g.pitch, dt, ntx, ntiles, ord.table, next_tickets(), dep_start_for(ord), current_trace());
}

// single-call host launchers (lagrange.cu, advec.cu)
void run_revert(const Grid& g, double* density0, double* density1, double* energy0, double* energy1);
void run_advec_mom(const Grid& g, int dirn, int sweep, double* vel_a, double* vel_b, double* mass_flux_x,
                   double* vol_flux_x, double* mass_flux_y, double* vol_flux_y, double* volume, double* density1,
                   double* celldx, double* celldy);
// advec_tma.cu
void run_advec_mom_tma(const Grid& g, int dirn, int sweep, double* vel_a, double* vel_b, double* mass_flux_x,
                       double* vol_flux_x, double* mass_flux_y, double* vol_flux_y, double* volume, double* density1,
                       double* celldx, double* celldy);

static bool same_grid(const Op& a, const Op& b) { return a.g.nx == b.g.nx && a.g.ny == b.g.ny; }

// ---- M: advec_mom pair ------------------------------------------------------------------------------------
static size_t fuse_mom_pair(const Op* q, size_t n, size_t i) {
  if (i + 1 >= n) return 0;
  const Op &x = q[i], &y = q[i + 1];
  if (x.kind != OP_ADVEC_MOM || y.kind != OP_ADVEC_MOM || !same_grid(x, y)) return 0;
  if (x.iv[0] != 1 || y.iv[0] != 2 || x.iv[1] != y.iv[1] || x.iv[2] != y.iv[2]) return 0;
  for (int k = 1; k < 9; ++k)
    if (x.a[k] != y.a[k]) return 0;
  if (x.a[0] == y.a[0]) return 0;
  if (tma_enabled()) {
    run_advec_mom_tma(x.g, x.iv[2], x.iv[1], x.a[0], y.a[0], x.a[1], x.a[2], x.a[3], x.a[4], x.a[5], x.a[6], x.a[7], x.a[8]);
    return 2;
  }
  // register/shuffle kernels: measured on B200 (profiles/) the two-component x kernel beats two launches (0.31 vs
  // 0.40 ms at 3840^2), the two-component y march does not (register pressure), so only x sweeps are paired
  if (x.iv[2] != 1) return 0;
  run_advec_mom(x.g, x.iv[2], x.iv[1], x.a[0], y.a[0], x.a[1], x.a[2], x.a[3], x.a[4], x.a[5], x.a[6], x.a[7], x.a[8]);
  return 2;
}

// ---- helpers for the halo ops inside a pattern ---------------------------------------------------------------
enum FieldId { F_DENSITY0 = 0, F_DENSITY1, F_ENERGY0, F_ENERGY1, F_PRESSURE, F_VISCOSITY, F_SOUNDSPEED, F_XVEL0,
               F_XVEL1, F_YVEL0, F_YVEL1 };
static bool halo_mask_is(const Op& o, std::initializer_list<int> ids, int depth) {
  if (o.iv[0] != depth) return false;
  int want[15] = {};
  for (int f : ids) want[f] = 1;
  for (int f = 0; f < 15; ++f)
    if ((o.fields[f] != 0) != (want[f] != 0)) return false;
  return true;
}
static HaloArgs halo_args(const Op& o, int drop_field) {
  HaloArgs h;
  for (int f = 0; f < 15; ++f) {
    h.host[f] = o.a[f];
    h.fields[f] = (f == drop_field) ? 0 : o.fields[f];
  }
  h.depth = o.iv[0];
  for (int f = 0; f < 4; ++f) h.ext[f] = o.iv[1 + f];
  return h;
}
// Optional [exchange][update_halo] pair with the given field set starting at q[i]; returns how many ops it is.
// `covered` reports whether all four faces of the chunk get their halo from these ops (neighbour faces from the
// exchange, external faces from update_halo).
static size_t match_halo(const Op* q, size_t n, size_t i, const Grid& g, std::initializer_list<int> ids, int depth,
                         const Op** ex, const Op** uh, bool* covered) {
  *ex = *uh = nullptr;
  size_t used = 0;
  if (i + used < n && q[i + used].kind == OP_EXCHANGE && halo_mask_is(q[i + used], ids, depth) &&
      q[i + used].g.nx == g.nx && q[i + used].g.ny == g.ny)
    *ex = &q[i + used++];
  if (i + used < n && q[i + used].kind == OP_UPDATE_HALO && halo_mask_is(q[i + used], ids, depth) &&
      q[i + used].g.nx == g.nx && q[i + used].g.ny == g.ny)
    *uh = &q[i + used++];
  if (covered) {
    int nb[4] = {-1, -1, -1, -1};
    if (chunk_registered())
      for (int f = 0; f < 4; ++f) nb[f] = chunk_neighbours()[f];
    bool all = true;
    for (int f = 0; f < 4; ++f) {
      const bool by_exchange = (*ex != nullptr) && nb[f] != -1;
      const bool by_reflect = (*uh != nullptr) && (*uh)->iv[1 + f] != 0;
      all = all && (by_exchange || by_reflect);
    }
    *covered = all;
  }
  return used;
}
// Is the value stored into `arr` by the op(s) before q[from] dead, i.e. fully overwritten before anything in the
// rest of the recorded stretch reads it?  (End of the stretch = the host may look = not dead.)
static bool dead_after(const Op* q, size_t n, size_t from, const double* arr) {
  for (size_t k = from; k < n; ++k) {
    if (q[k].does_read(arr)) return false;
    if (q[k].does_overwrite(arr)) return true;
    if (q[k].touches(arr)) return false;
  }
  return false;
}

static int g_ctas_per_sm_timestep[2] = {0, 0};

// the viscosity halo update of the timestep pattern, waiting to be merged into the next pressure halo update
struct PendingHalo {
  bool active = false, has_ex = false, has_uh = false;
  Grid g{};
  HaloArgs ex{}, uh{};
};
static PendingHalo g_pending_halo;
static bool merge_halo_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("CLOVER_B200_MERGE_HALO") ? atoi(getenv("CLOVER_B200_MERGE_HALO")) : 1;
  return v != 0;
}
bool pending_halo_exists() { return g_pending_halo.active; }
void run_pending_halo() {
  PendingHalo& P = g_pending_halo;
  if (!P.active) return;
  P.active = false;
  run_exchange_then_halo(P.g, P.has_ex ? &P.ex : nullptr, P.has_uh ? &P.uh : nullptr);
}

// ---- T: ideal_gas -> halo{d0,e0,p,u0,v0} -> viscosity -> halo{q} -> calc_dt -------------------------------------
static size_t fuse_timestep(const Op* q, size_t n, size_t i) {
  const Op& ig = q[i];
  if (ig.kind != OP_IDEAL_GAS) return 0;
  const Grid g = ig.g;
  size_t k = i + 1;
  const Op *ex1, *uh1, *ex2, *uh2;
  bool covered = false;
  k += match_halo(q, n, k, g, {F_DENSITY0, F_ENERGY0, F_PRESSURE, F_XVEL0, F_YVEL0}, 1, &ex1, &uh1, &covered);
  if (!covered || k >= n || q[k].kind != OP_VISCOSITY || !same_grid(ig, q[k])) return 0;
  const Op& vi = q[k++];
  k += match_halo(q, n, k, g, {F_VISCOSITY}, 1, &ex2, &uh2, nullptr);
  if (k >= n || q[k].kind != OP_CALC_DT || !same_grid(ig, q[k])) return 0;
  const Op& dt = q[k++];
  // the same arrays all the way through
  double *density0 = ig.a[0], *energy0 = ig.a[1], *pressure = ig.a[2], *soundspeed = ig.a[3];
  double *celldx = vi.a[0], *celldy = vi.a[1], *viscosity = vi.a[4], *xvel0 = vi.a[5], *yvel0 = vi.a[6];
  if (vi.a[2] != density0 || vi.a[3] != pressure) return 0;
  if (dt.a[2] != celldx || dt.a[3] != celldy || dt.a[5] != density0 || dt.a[6] != viscosity || dt.a[7] != soundspeed ||
      dt.a[8] != xvel0 || dt.a[9] != yvel0)
    return 0;
  for (const Op* h : {ex1, uh1})
    if (h && (h->a[F_DENSITY0] != density0 || h->a[F_ENERGY0] != energy0 || h->a[F_PRESSURE] != pressure ||
              h->a[F_XVEL0] != xvel0 || h->a[F_YVEL0] != yvel0))
      return 0;
  for (const Op* h : {ex2, uh2})
    if (h && h->a[F_VISCOSITY] != viscosity) return 0;
  double *xarea = dt.a[0], *yarea = dt.a[1], *volume = dt.a[4];

  // halos of density0, energy0, xvel0, yvel0 first (pressure's ring is produced by the kernel itself)
  {
    HaloArgs hx, hu;
    if (ex1) hx = halo_args(*ex1, F_PRESSURE);
    if (uh1) hu = halo_args(*uh1, F_PRESSURE);
    run_exchange_then_halo(g, ex1 ? &hx : nullptr, uh1 ? &hu : nullptr);
  }
  {
    const double* xa = dev(g, xarea, XFACE, IN);
    const double* ya = dev(g, yarea, YFACE, IN);
    const double* cdx = dev(g, celldx, X1D_CELL, IN);
    const double* cdy = dev(g, celldy, Y1D_CELL, IN);
    const double* vol = dev(g, volume, CELL, IN);
    const double* d0 = dev(g, density0, CELL, IN);
    const double* e0 = dev(g, energy0, CELL, IN);
    double* p = dev(g, pressure, CELL, OUT_FULL);
    double* qv = dev(g, viscosity, CELL, OUT_FULL);
    // the sound speed is consumed on chip; in resident mode the array is only marked "c(density0, energy0), not
    // evaluated yet" (runtime.cu: lazy_soundspeed) -- one store pass less
    const bool lazy_ss = is_resident() && tma_enabled();
    double* ss = dev(g, soundspeed, CELL, OUT_FULL);
    if (lazy_ss) lazy_soundspeed(g, soundspeed, density0, energy0);
    const double* xv = dev(g, xvel0, VERTEX, IN);
    const double* yv = dev(g, yvel0, VERTEX, IN);
    const DtParams P{dt.sv[0], dt.sv[1], dt.sv[2], dt.sv[3], dt.sv[4], dt.sv[5]};
    const Range r = make_range(1, g.nx, 1, g.ny);
    const ReduceTail RT = next_reduce_tail(0);
    set_dt_result_seq(RT.seq);
    if (!g_ctas_per_sm_timestep[0]) {
      CLV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_ctas_per_sm_timestep[0], timestep_kernel<true>,
                                                             BX * BY, 0));
      if (g_ctas_per_sm_timestep[0] < 1) g_ctas_per_sm_timestep[0] = 1;
    }
    if (tma_enabled()) {
      static bool configured = false;
      if (!configured) {
        CLV_CUDA(cudaFuncSetAttribute(timestep_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TT_SMEM));
        CLV_CUDA(cudaFuncSetAttribute(timestep_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TT_SMEM));
        configured = true;
      }
      TimestepMaps M;
      const double* in[TT_NARR] = {d0, e0, xv, yv, vol, xa, ya};
      for (int a = 0; a < TT_NARR; ++a) M.m[a] = *tensor_map_for(g, in[a], TT_BW, TT_BH);
      const int ntx = (g.nx + TT_W - 1) / TT_W, nty = (g.ny + TT_H - 1) / TT_H;
      const int ntiles = ntx * nty;
      const int cap = sm_count() * TT_CPS;
      const int ctas = ntiles < cap ? ntiles : cap;
      double* part = partials((size_t)ctas);
      LaunchScope ls("timestep_tma");
      const TileOrder ord = tile_order_split(ntx, nty, TT_W, TT_H, 2, TT_BW - 2 - TT_W, 1, TT_BH - 1 - TT_H, g.nx, g.ny);
      launch_pdl(lazy_ss ? timestep_tma_kernel<false> : timestep_tma_kernel<true>, dim3(ctas), dim3(BX, BY), TT_SMEM, stream(), M,
                 P, cdx, cdy, d0, e0, p, qv, ss, part, ticket(), host_scalars(), g.nx, g.ny, g.pitch, ntx, ntiles, ord.table,
                 next_tickets(), dep_start_for(ord), ls.trace, RT);
    } else {
    const dim3 grid = persistent_grid(r, 1, g_ctas_per_sm_timestep[0]);
    double* part = partials((size_t)grid.x * grid.y);
    LaunchScope ls("timestep_fused");
    timestep_kernel<true><<<grid, dim3(BX, BY), 0, stream()>>>(r, g.pitch, P, xa, ya, cdx, cdy, vol, d0, e0, p, qv, ss,
                                                               xv, yv, part, ticket(), host_scalars(), RT);
    }
    note_fused_allreduce(0, 1, true, RT.all != nullptr);
  }
  if (ex2 || uh2) {
    // Nothing reads the viscosity halo before accelerate (SURVEY 8a'): its exchange + reflective boundary are not
    // issued here but kept pending, and ride along with the pressure exchange of the PdV predictor pattern, one
    // NVLink round trip later in the step instead of two (fuse_predict below).  Anything else that comes first --
    // another call sequence, a download, the end of the run -- issues them on their own (runtime.cu: flush_deferred).
    PendingHalo& P = g_pending_halo;
    if (P.active) run_pending_halo();
    P.active = true;
    P.g = g;
    P.has_ex = ex2 != nullptr;
    P.has_uh = uh2 != nullptr;
    if (ex2) P.ex = halo_args(*ex2, -1);
    if (uh2) P.uh = halo_args(*uh2, -1);
    if (!(is_resident() && fusion_enabled() && merge_halo_enabled())) run_pending_halo();
  }
  return k - i;
}

// ---- P: PdV predictor -> ideal_gas(d1,e1) -> halo{p} -> revert ---------------------------------------------------
static size_t fuse_predict(const Op* q, size_t n, size_t i) {
  const Op& pv = q[i];
  if (pv.kind != OP_PDV_PREDICT || i + 2 >= n) return 0;
  const Op& ig = q[i + 1];
  if (ig.kind != OP_IDEAL_GAS || !same_grid(pv, ig)) return 0;
  const Grid g = pv.g;
  double *xarea = pv.a[0], *yarea = pv.a[1], *volume = pv.a[2], *density0 = pv.a[3], *density1 = pv.a[4],
         *energy0 = pv.a[5], *energy1 = pv.a[6], *pressure = pv.a[7], *viscosity = pv.a[8], *xvel0 = pv.a[9],
         *yvel0 = pv.a[11];
  if (ig.a[0] != density1 || ig.a[1] != energy1 || ig.a[2] != pressure) return 0;
  double* soundspeed = ig.a[3];
  size_t k = i + 2;
  const Op *ex, *uh;
  k += match_halo(q, n, k, g, {F_PRESSURE}, 1, &ex, &uh, nullptr);
  for (const Op* h : {ex, uh})
    if (h && h->a[F_PRESSURE] != pressure) return 0;
  if (k >= n || q[k].kind != OP_REVERT || !same_grid(pv, q[k])) return 0;
  const Op& rv = q[k++];
  if (rv.a[0] != density0 || rv.a[1] != density1 || rv.a[2] != energy0 || rv.a[3] != energy1) return 0;
  const bool write_ss = !dead_after(q, n, i + 2, soundspeed);
  {
    const double* xa = dev(g, xarea, XFACE, IN);
    const double* ya = dev(g, yarea, YFACE, IN);
    const double* vol = dev(g, volume, CELL, IN);
    const double* d0 = dev(g, density0, CELL, IN);
    const double* e0 = dev(g, energy0, CELL, IN);
    double* p = dev(g, pressure, CELL, INOUT);
    const double* qv = dev(g, viscosity, CELL, IN);
    // written in full, or dead (overwritten later in this stretch before anything reads it): either way a pending
    // lazy evaluation of the sound speed (fuse_timestep) is superseded
    double* ss = dev(g, soundspeed, CELL, OUT_FULL);
    const double* x0 = dev(g, xvel0, VERTEX, IN);
    const double* y0 = dev(g, yvel0, VERTEX, IN);
    if (tma_enabled()) {
      static bool configured = false;
      if (!configured) {
        CLV_CUDA(cudaFuncSetAttribute(pdv_predict_eos_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SMEM));
        CLV_CUDA(cudaFuncSetAttribute(pdv_predict_eos_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SMEM));
        configured = true;
      }
      PredictMaps M;
      const double* in[PT_NARR] = {xa, ya, vol, d0, e0, p, qv, x0, y0};
      for (int a = 0; a < PT_NARR; ++a) M.m[a] = *tensor_map_for(g, in[a], PT_BW, PT_BH);
      const int ntx = (g.nx + PT_W - 1) / PT_W, nty = (g.ny + PT_H - 1) / PT_H;
      const int ntiles = ntx * nty;
      const int cap = sm_count() * PT_CPS;
      const int ctas = ntiles < cap ? ntiles : cap;
      LaunchScope ls("pdv_predict_tma");
      const TileOrder ord = tile_order_split(ntx, nty, PT_W, PT_H, 0, PT_BW - PT_W, 0, PT_BH - PT_H, g.nx, g.ny);
      const int dep = dep_start_for(ord);
      const Tickets tk = next_tickets();
      if (write_ss)
        launch_pdl(pdv_predict_eos_tma_kernel<true>, dim3(ctas), dim3(BX, BY), PT_SMEM, stream(), M, pv.sv[0], p, ss, g.nx, g.ny,
                   g.pitch, ntx, ntiles, ord.table, tk, dep, ls.trace);
      else
        launch_pdl(pdv_predict_eos_tma_kernel<false>, dim3(ctas), dim3(BX, BY), PT_SMEM, stream(), M, pv.sv[0], p, ss, g.nx, g.ny,
                   g.pitch, ntx, ntiles, ord.table, tk, dep, ls.trace);
    } else {
    const Range r = make_range(1, g.nx, 1, g.ny);
    const dim3 grid = grid_for(r, NR_PRED);
    LaunchScope ls("pdv_predict_fused");
    if (write_ss)
      pdv_predict_eos_kernel<true><<<grid, dim3(BX, BY), 0, stream()>>>(r, g.pitch, pv.sv[0], xa, ya, vol, d0, e0, p, qv,
                                                                        ss, x0, y0);
    else
      pdv_predict_eos_kernel<false><<<grid, dim3(BX, BY), 0, stream()>>>(r, g.pitch, pv.sv[0], xa, ya, vol, d0, e0, p,
                                                                         qv, ss, x0, y0);
    }
  }
  join_side();
  {
    HaloArgs hx, hu;
    if (ex) hx = halo_args(*ex, -1);
    if (uh) hu = halo_args(*uh, -1);
    // the pending viscosity halo update (fuse_timestep) rides along: same depth, same chunk, one more cell field
    PendingHalo& P = g_pending_halo;
    if (P.active && P.g.nx == g.nx && P.g.ny == g.ny && P.has_ex == (ex != nullptr) && P.has_uh == (uh != nullptr) &&
        (!ex || P.ex.depth == hx.depth) && (!uh || (P.uh.depth == hu.depth && memcmp(P.uh.ext, hu.ext, sizeof(hu.ext)) == 0))) {
      for (int f = 0; f < 15; ++f) {
        if (ex && P.ex.fields[f]) { hx.fields[f] = 1; hx.host[f] = P.ex.host[f]; }
        if (uh && P.uh.fields[f]) { hu.fields[f] = 1; hu.host[f] = P.uh.host[f]; }
      }
      P.active = false;
    } else {
      run_pending_halo();
    }
    run_exchange_then_halo(g, ex ? &hx : nullptr, uh ? &hu : nullptr);
  }
  run_revert(g, density0, density1, energy0, energy1);  // a lazy copy in resident mode
  return k - i;
}

// ---- C: accelerate -> PdV corrector -> flux_calc -------------------------------------------------------------------
static size_t fuse_correct(const Op* q, size_t n, size_t i) {
  if (i + 2 >= n) return 0;
  const Op &ac = q[i], &pv = q[i + 1], &fc = q[i + 2];
  if (ac.kind != OP_ACCELERATE || pv.kind != OP_PDV_CORRECT || fc.kind != OP_FLUX_CALC) return 0;
  if (!same_grid(ac, pv) || !same_grid(ac, fc) || ac.sv[0] != pv.sv[0] || ac.sv[0] != fc.sv[0]) return 0;
  const Grid g = ac.g;
  double *xarea = ac.a[0], *yarea = ac.a[1], *volume = ac.a[2], *density0 = ac.a[3], *pressure = ac.a[4],
         *viscosity = ac.a[5], *xvel0 = ac.a[6], *yvel0 = ac.a[7], *xvel1 = ac.a[8], *yvel1 = ac.a[9];
  double *density1 = pv.a[4], *energy0 = pv.a[5], *energy1 = pv.a[6];
  if (pv.a[0] != xarea || pv.a[1] != yarea || pv.a[2] != volume || pv.a[3] != density0 || pv.a[7] != pressure ||
      pv.a[8] != viscosity || pv.a[9] != xvel0 || pv.a[10] != xvel1 || pv.a[11] != yvel0 || pv.a[12] != yvel1)
    return 0;
  if (fc.a[0] != xarea || fc.a[1] != yarea || fc.a[2] != xvel0 || fc.a[3] != yvel0 || fc.a[4] != xvel1 ||
      fc.a[5] != yvel1)
    return 0;
  double *vol_flux_x = fc.a[6], *vol_flux_y = fc.a[7];
  CorrectArgs A;
  A.xarea = dev(g, xarea, XFACE, IN);
  A.yarea = dev(g, yarea, YFACE, IN);
  A.volume = dev(g, volume, CELL, IN);
  A.density0 = dev(g, density0, CELL, IN);
  A.energy0 = dev(g, energy0, CELL, IN);
  A.pressure = dev(g, pressure, CELL, IN);
  A.viscosity = dev(g, viscosity, CELL, IN);
  A.xvel0 = dev(g, xvel0, VERTEX, IN);
  A.yvel0 = dev(g, yvel0, VERTEX, IN);
  A.xvel1 = dev(g, xvel1, VERTEX, OUT_FULL);
  A.yvel1 = dev(g, yvel1, VERTEX, OUT_FULL);
  A.density1 = dev(g, density1, CELL, OUT_FULL);
  A.energy1 = dev(g, energy1, CELL, OUT_FULL);
  A.vol_flux_x = dev(g, vol_flux_x, XFACE, OUT_FULL);
  A.vol_flux_y = dev(g, vol_flux_y, YFACE, OUT_FULL);
  if (tma_enabled()) {
    // <tile width, rows per thread, ring stages, CTAs per SM>; measured on B200 at 3840^2: <64,2,2,2> 0.337 ms,
    // <64,1,4,1> 0.424, <64,2,4,1> 0.452, <64,4,2,2> 0.369, <64,1,2,2> 0.403, <32,2,2,3> 0.358, <32,2,2,4> 0.430
    LaunchScope ls("lagrange_correct_tma");
    launch_correct_tma<LC_W, 2, 2, LC_CPS>(A, g, ac.sv[0]);
    return 3;
  }
  const dim3 grid((unsigned)((g.nx + 1 + CT_W - 1) / CT_W), (unsigned)((g.ny + 1 + CT_H - 1) / CT_H));
  LaunchScope ls("lagrange_correct_fused");
  lagrange_correct_kernel<<<grid, dim3(CT_W, CT_BY), 0, stream()>>>(A, g.nx, g.ny, g.pitch, ac.sv[0]);
  return 3;
}

// ---- X: clover_exchange -> update_halo_kernel with the same field list -------------------------------------------------
static size_t fuse_exchange_halo(const Op* q, size_t n, size_t i) {
  if (i + 1 >= n || q[i].kind != OP_EXCHANGE || q[i + 1].kind != OP_UPDATE_HALO || !same_grid(q[i], q[i + 1])) return 0;
  const HaloArgs hx = halo_args(q[i], -1), hu = halo_args(q[i + 1], -1);
  run_exchange_then_halo(q[i].g, &hx, &hu);
  return 2;
}

size_t fuse_at(const Op* q, size_t n, size_t i) {
  switch (q[i].kind) {
    case OP_EXCHANGE: return fuse_exchange_halo(q, n, i);
    case OP_ADVEC_MOM: return fuse_mom_pair(q, n, i);
    case OP_IDEAL_GAS: return fuse_timestep(q, n, i);
    case OP_PDV_PREDICT: return fuse_predict(q, n, i);
    case OP_ACCELERATE: return fuse_correct(q, n, i);
    default: return 0;
  }
}

}  // namespace clv
