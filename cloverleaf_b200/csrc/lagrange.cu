// lagrange.cu -- the Lagrangian-phase kernels and the two reductions, fp64 CUDA for sm_100a:
//   ideal_gas, viscosity, calc_dt (min-reduction), PdV predictor/corrector, revert, accelerate,
//   flux_calc, reset_field, field_summary (sum-reductions).
//
// Numerics: built with -fmad=false; every expression keeps the evaluation order of the reference's
// C source (cited per kernel) so fields are bit-identical to CloverLeaf_ref's C kernels compiled
// with -ffp-contract=off.  Reductions: min is order-independent; sums use a fixed tree
// (deterministic for a given mesh, differs from a serial CPU sum in the last bits only).
//
// Mapping: threadIdx.x along j (unit stride, 128-byte aligned rows, see common.cuh), 32x8 thread
// tiles, and every thread handles NR rows (k, k+8, ...) in a fully unrolled loop whose loads are
// unconditional (indices clamped into the loop range, only the stores are predicated).  That puts
// NR independent load batches and NR independent fp64 dependency chains (the div/sqrt sequences are
// ~10 dependent DFMAs each) in flight per thread: first profiles (profiles/r01b_full_ddiv.txt) showed
// these kernels neither DRAM- nor issue-bound but latency-bound with one cell per thread.
// Stencil neighbours come through L1.  Algorithmic bytes per cell are listed in DESIGN.md.
#include "clover_b200.h"
#include "common.cuh"
#include "lagrange.cuh"

namespace clv {

// the sound speed alone (a pending lazy_soundspeed that something does read after all; runtime.cu)
__global__ void __launch_bounds__(BX* BY)
    soundspeed_kernel(Range r, int pitch, const double* __restrict__ density, const double* __restrict__ energy,
                      double* __restrict__ soundspeed) {
  CLV_ROWS_BEGIN(r, 1)
    const size_t c = idx2(pitch, j, k);
    const double rho = density[c], en = energy[c];
    bool bad = false;
    double p, ss;
    ideal_gas_cell<false>(rho, en, p, ss, bad);
    if (bad) ideal_gas_cell<true>(rho, en, p, ss, bad);
    if (active) soundspeed[c] = ss;
  CLV_ROWS_END
}

template <int NR>
__global__ void __launch_bounds__(BX* BY)
    ideal_gas_kernel(Range r, int pitch, const double* __restrict__ density,
                     const double* __restrict__ energy, double* __restrict__ pressure,
                     double* __restrict__ soundspeed) {
  double rho[NR], en[NR];
  CLV_ROWS_BEGIN(r, NR)  // all loads first
    const size_t c = idx2(pitch, j, k);
    rho[rr_] = density[c];
    en[rr_] = energy[c];
    if (k + PF_ROWS <= r.k1) {
      prefetch_l2(density + c + (size_t)PF_ROWS * pitch);
      prefetch_l2(energy + c + (size_t)PF_ROWS * pitch);
    }
    (void)active;
  CLV_ROWS_END
  double p[NR], ss[NR];
  bool bad = false;
#pragma unroll
  for (int i = 0; i < NR; ++i) ideal_gas_cell<false>(rho[i], en[i], p[i], ss[i], bad);
  if (bad) {
#pragma unroll
    for (int i = 0; i < NR; ++i) ideal_gas_cell<true>(rho[i], en[i], p[i], ss[i], bad);
  }
  {
    CLV_ROWS_BEGIN(r, NR)
      if (active) {
        const size_t c = idx2(pitch, j, k);
        pressure[c] = p[rr_];
        soundspeed[c] = ss[rr_];
      }
    CLV_ROWS_END
  }
}

template <int NR>
__global__ void __launch_bounds__(BX* BY)
    viscosity_kernel(Range r, int pitch, const double* __restrict__ celldx,
                     const double* __restrict__ celldy, const double* __restrict__ density0,
                     const double* __restrict__ pressure, double* __restrict__ viscosity,
                     const double* __restrict__ xvel0, const double* __restrict__ yvel0) {
  ViscIn I[NR];
  CLV_ROWS_BEGIN(r, NR)  // phase 1: every load of every row, back to back
    const size_t c = idx2(pitch, j, k);
    ViscIn& in = I[rr_];
    in.u00 = xvel0[c]; in.u10 = xvel0[c + 1]; in.u01 = xvel0[c + pitch]; in.u11 = xvel0[c + pitch + 1];
    in.v00 = yvel0[c]; in.v10 = yvel0[c + 1]; in.v01 = yvel0[c + pitch]; in.v11 = yvel0[c + pitch + 1];
    in.dx = celldx[j + 1]; in.dy = celldy[k + 1]; in.dx1 = celldx[j + 2]; in.dy1 = celldy[k + 2];
    in.pl = pressure[c - 1]; in.pr = pressure[c + 1]; in.pb = pressure[c - pitch]; in.pt = pressure[c + pitch];
    in.rho = density0[c];
    if (k + PF_ROWS <= r.k1) {
      const size_t pf = c + (size_t)PF_ROWS * pitch;
      prefetch_l2(xvel0 + pf);
      prefetch_l2(yvel0 + pf);
      prefetch_l2(pressure + pf);
      prefetch_l2(density0 + pf);
    }
    (void)active;
  CLV_ROWS_END
  double q[NR];
  bool bad = false;
#pragma unroll
  for (int i = 0; i < NR; ++i) q[i] = viscosity_cell<false>(I[i], bad);
  if (bad) {
#pragma unroll
    for (int i = 0; i < NR; ++i) q[i] = viscosity_cell<true>(I[i], bad);
  }
  {
    CLV_ROWS_BEGIN(r, NR)
      if (active) viscosity[idx2(pitch, j, k)] = q[rr_];
    CLV_ROWS_END
  }
}

template <int NR>
__global__ void __launch_bounds__(BX* BY)
    calc_dt_kernel(Range r, int pitch, DtParams P, const double* __restrict__ xarea,
                   const double* __restrict__ yarea, const double* __restrict__ celldx,
                   const double* __restrict__ celldy, const double* __restrict__ volume,
                   const double* __restrict__ density0, const double* __restrict__ viscosity,
                   const double* __restrict__ soundspeed, const double* __restrict__ xvel0,
                   const double* __restrict__ yvel0, double* __restrict__ partials,
                   unsigned int* ticket, double* __restrict__ out, ReduceTail RT) {
  double m[1] = {P.g_big};
  CLV_PTILES_BEGIN(r, NR)
    const size_t c = idx2(pitch, j, k);
    // all loads first (one batch in flight), then the arithmetic
    const double dsx = celldx[j + 1], dsy = celldy[k + 1];
    const double vol = volume[c], ssp = soundspeed[c], visc = viscosity[c], rho = density0[c];
    const double u00 = xvel0[c], u10 = xvel0[c + 1], u01 = xvel0[c + pitch], u11 = xvel0[c + pitch + 1];
    const double v00 = yvel0[c], v10 = yvel0[c + 1], v01 = yvel0[c + pitch], v11 = yvel0[c + pitch + 1];
    const double xa0 = xarea[c], xa1 = xarea[c + 1], ya0 = yarea[c], ya1 = yarea[c + pitch];
    if (k + PF_ROWS <= r.k1) {
      const size_t pf = c + (size_t)PF_ROWS * pitch;
      prefetch_l2(volume + pf); prefetch_l2(soundspeed + pf); prefetch_l2(viscosity + pf); prefetch_l2(density0 + pf);
      prefetch_l2(xvel0 + pf); prefetch_l2(yvel0 + pf); prefetch_l2(xarea + pf); prefetch_l2(yarea + pf);
    }
    DtIn in{dsx, dsy, vol, ssp, visc, rho, u00, u10, u01, u11, v00, v10, v01, v11, xa0, xa1, ya0, ya1};
    bool bad = false;
    double cell_dt = calc_dt_cell<false>(in, P, bad);
    if (bad) cell_dt = calc_dt_cell<true>(in, P, bad);
    if (active && cell_dt < m[0]) m[0] = cell_dt;
  CLV_PTILES_END
  block_reduce_publish<1, true>(m, partials, ticket, out, P.g_big, RT);
}

// field_summary_kernel_c.c:66-89.  6 passes read = 48 B/cell.
template <int NR>
__global__ void __launch_bounds__(BX* BY)
    field_summary_kernel(Range r, int pitch, const double* __restrict__ volume,
                         const double* __restrict__ density0, const double* __restrict__ energy0,
                         const double* __restrict__ pressure, const double* __restrict__ xvel0,
                         const double* __restrict__ yvel0, double* __restrict__ partials,
                         unsigned int* ticket, double* __restrict__ out, ReduceTail RT) {
  double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};  // vol, mass, ie, ke, press
  CLV_PTILES_BEGIN(r, NR)
    const size_t c = idx2(pitch, j, k);
    double vsqrd = 0.0;
    vsqrd = vsqrd + 0.25 * (xvel0[c] * xvel0[c] + yvel0[c] * yvel0[c]);
    vsqrd = vsqrd + 0.25 * (xvel0[c + 1] * xvel0[c + 1] + yvel0[c + 1] * yvel0[c + 1]);
    vsqrd = vsqrd + 0.25 * (xvel0[c + pitch] * xvel0[c + pitch] + yvel0[c + pitch] * yvel0[c + pitch]);
    vsqrd = vsqrd + 0.25 * (xvel0[c + pitch + 1] * xvel0[c + pitch + 1] + yvel0[c + pitch + 1] * yvel0[c + pitch + 1]);
    const double cell_vol = volume[c];
    const double cell_mass = cell_vol * density0[c];
    if (active) {
      s[0] += cell_vol;
      s[1] += cell_mass;
      s[2] += cell_mass * energy0[c];
      s[3] += cell_mass * 0.5 * vsqrd;
      s[4] += cell_vol * pressure[c];
    }
  CLV_PTILES_END
  block_reduce_publish<5, false>(s, partials, ticket, out, 0.0, RT);
}

// ------------------------------------------------------------------------------------------------
// PdV_kernel_c.c:63-113 (predictor) / :115-167 (corrector).  11 / 13 passes.
template <bool PREDICT, int NR>
__global__ void __launch_bounds__(BX* BY)
    pdv_kernel(Range r, int pitch, double dt, const double* __restrict__ xarea,
               const double* __restrict__ yarea, const double* __restrict__ volume,
               const double* __restrict__ density0, double* __restrict__ density1,
               const double* __restrict__ energy0, double* __restrict__ energy1,
               const double* __restrict__ pressure, const double* __restrict__ viscosity,
               const double* __restrict__ xvel0, const double* __restrict__ xvel1,
               const double* __restrict__ yvel0, const double* __restrict__ yvel1) {
  CLV_ROWS_BEGIN(r, NR)
    const size_t c = idx2(pitch, j, k);
    const double x00 = xvel0[c], x10 = xvel0[c + 1], x01 = xvel0[c + pitch], x11 = xvel0[c + pitch + 1];
    const double y00 = yvel0[c], y10 = yvel0[c + 1], y01 = yvel0[c + pitch], y11 = yvel0[c + pitch + 1];
    const double vol = volume[c], rho0 = density0[c], pres = pressure[c], visc = viscosity[c], en0 = energy0[c];
    if (k + PF_ROWS <= r.k1) {
      const size_t pf = c + (size_t)PF_ROWS * pitch;
      prefetch_l2(xvel0 + pf); prefetch_l2(yvel0 + pf); prefetch_l2(volume + pf); prefetch_l2(density0 + pf);
      prefetch_l2(pressure + pf); prefetch_l2(viscosity + pf); prefetch_l2(energy0 + pf); prefetch_l2(xarea + pf);
      prefetch_l2(yarea + pf);
      if (!PREDICT) { prefetch_l2(xvel1 + pf); prefetch_l2(yvel1 + pf); }
    }
    double left, right, bottom, top;
    if (PREDICT) {
      left = xarea[c] * (x00 + x01 + x00 + x01) * 0.25 * dt * 0.5;
      right = xarea[c + 1] * (x10 + x11 + x10 + x11) * 0.25 * dt * 0.5;
      bottom = yarea[c] * (y00 + y10 + y00 + y10) * 0.25 * dt * 0.5;
      top = yarea[c + pitch] * (y01 + y11 + y01 + y11) * 0.25 * dt * 0.5;
    } else {
      const double a00 = xvel1[c], a10 = xvel1[c + 1], a01 = xvel1[c + pitch], a11 = xvel1[c + pitch + 1];
      const double b00 = yvel1[c], b10 = yvel1[c + 1], b01 = yvel1[c + pitch], b11 = yvel1[c + pitch + 1];
      left = xarea[c] * (x00 + x01 + a00 + a01) * 0.25 * dt;
      right = xarea[c + 1] * (x10 + x11 + a10 + a11) * 0.25 * dt;
      bottom = yarea[c] * (y00 + y10 + b00 + b10) * 0.25 * dt;
      top = yarea[c + pitch] * (y01 + y11 + b01 + b11) * 0.25 * dt;
    }
    const double total = right - left + top - bottom;
    const double vc = vol / (vol + total);
    const double recip = 1.0 / vol;
    const double de = (pres / rho0 + ddiv(visc, rho0)) * total * recip;
    const double e1 = en0 - de;
    if (active) {
      energy1[c] = e1;
      density1[c] = rho0 * vc;
    }
  CLV_ROWS_END
}

// revert_kernel_c.c:46-62 (2 copies) and reset_field_kernel_c.c:46-76 (4 copies)
template <int NR>
__global__ void __launch_bounds__(BX* BY)
    copy2_kernel(Range r, int pitch, const double* __restrict__ a_src, double* __restrict__ a_dst,
                 const double* __restrict__ b_src, double* __restrict__ b_dst) {
  CLV_ROWS_BEGIN(r, NR)
    const size_t c = idx2(pitch, j, k);
    const double a = a_src[c], b = b_src[c];
    if (k + PF_ROWS <= r.k1) {
      prefetch_l2(a_src + c + (size_t)PF_ROWS * pitch);
      prefetch_l2(b_src + c + (size_t)PF_ROWS * pitch);
    }
    if (active) {
      a_dst[c] = a;
      b_dst[c] = b;
    }
  CLV_ROWS_END
}
template <int NR>
__global__ void __launch_bounds__(BX* BY)
    reset_field_kernel(Range r, int pitch, int nx, int ny, double* __restrict__ density0,
                       const double* __restrict__ density1, double* __restrict__ energy0,
                       const double* __restrict__ energy1, double* __restrict__ xvel0,
                       const double* __restrict__ xvel1, double* __restrict__ yvel0,
                       const double* __restrict__ yvel1) {
  CLV_ROWS_BEGIN(r, NR)
    const size_t c = idx2(pitch, j, k);
    const double d = density1[c], e = energy1[c], u = xvel1[c], v = yvel1[c];
    if (k + PF_ROWS <= r.k1) {
      const size_t pf = c + (size_t)PF_ROWS * pitch;
      prefetch_l2(density1 + pf); prefetch_l2(energy1 + pf); prefetch_l2(xvel1 + pf); prefetch_l2(yvel1 + pf);
    }
    if (active) {
      if (j <= nx && k <= ny) {
        density0[c] = d;
        energy0[c] = e;
      }
      xvel0[c] = u;
      yvel0[c] = v;
    }
  CLV_ROWS_END
}

// accelerate_kernel_c.c:56-95.  10 passes = 80 B/cell.
template <int NR>
__global__ void __launch_bounds__(BX* BY)
    accelerate_kernel(Range r, int pitch, double dt, const double* __restrict__ xarea,
                      const double* __restrict__ yarea, const double* __restrict__ volume,
                      const double* __restrict__ density0, const double* __restrict__ pressure,
                      const double* __restrict__ viscosity, const double* __restrict__ xvel0,
                      const double* __restrict__ yvel0, double* __restrict__ xvel1,
                      double* __restrict__ yvel1) {
  CLV_ROWS_BEGIN(r, NR)
    const size_t c11 = idx2(pitch, j, k), c01 = c11 - 1, c10 = c11 - pitch, c00 = c10 - 1;
    // all loads first (one batch in flight), then the arithmetic
    const double d00 = density0[c00], d10 = density0[c10], d11 = density0[c11], d01 = density0[c01];
    const double w00 = volume[c00], w10 = volume[c10], w11 = volume[c11], w01 = volume[c01];
    const double xa1 = xarea[c11], xa0 = xarea[c10];
    const double ya1 = yarea[c11], ya0 = yarea[c01];
    const double p11 = pressure[c11], p01 = pressure[c01], p10 = pressure[c10], p00 = pressure[c00];
    const double q11 = viscosity[c11], q01 = viscosity[c01], q10 = viscosity[c10], q00 = viscosity[c00];
    const double xv0 = xvel0[c11], yv0 = yvel0[c11];
    if (k + PF_ROWS <= r.k1) {
      const size_t pf = c11 + (size_t)PF_ROWS * pitch;
      prefetch_l2(density0 + pf); prefetch_l2(volume + pf); prefetch_l2(xarea + pf); prefetch_l2(yarea + pf);
      prefetch_l2(pressure + pf); prefetch_l2(viscosity + pf); prefetch_l2(xvel0 + pf); prefetch_l2(yvel0 + pf);
    }
    const double nodal_mass = (d00 * w00 + d10 * w10 + d11 * w11 + d01 * w01) * 0.25;
    const double s = 0.5 * dt / nodal_mass;
    double xv = xv0 - s * (xa1 * (p11 - p01) + xa0 * (p10 - p00));
    double yv = yv0 - s * (ya1 * (p11 - p10) + ya0 * (p01 - p00));
    xv = xv - s * (xa1 * (q11 - q01) + xa0 * (q10 - q00));
    yv = yv - s * (ya1 * (q11 - q10) + ya0 * (q01 - q00));
    if (active) {
      xvel1[c11] = xv;
      yvel1[c11] = yv;
    }
  CLV_ROWS_END
}

// flux_calc_kernel_c.c:49-73.  8 passes = 64 B/cell.
template <int NR>
__global__ void __launch_bounds__(BX* BY)
    flux_calc_kernel(Range r, int pitch, int nx, int ny, double dt, const double* __restrict__ xarea,
                     const double* __restrict__ yarea, const double* __restrict__ xvel0,
                     const double* __restrict__ yvel0, const double* __restrict__ xvel1,
                     const double* __restrict__ yvel1, double* __restrict__ vol_flux_x,
                     double* __restrict__ vol_flux_y) {
  CLV_ROWS_BEGIN(r, NR)
    const size_t c = idx2(pitch, j, k);
    const double x0 = xvel0[c], x1 = xvel1[c], y0 = yvel0[c], y1 = yvel1[c];
    if (k + PF_ROWS <= r.k1) {
      const size_t pf = c + (size_t)PF_ROWS * pitch;
      prefetch_l2(xvel0 + pf); prefetch_l2(xvel1 + pf); prefetch_l2(yvel0 + pf); prefetch_l2(yvel1 + pf);
      prefetch_l2(xarea + pf); prefetch_l2(yarea + pf);
    }
    const double fx = 0.25 * dt * xarea[c] * (x0 + xvel0[c + pitch] + x1 + xvel1[c + pitch]);
    const double fy = 0.25 * dt * yarea[c] * (y0 + yvel0[c + 1] + y1 + yvel1[c + 1]);
    if (active) {
      if (k <= ny) vol_flux_x[c] = fx;
      if (j <= nx) vol_flux_y[c] = fy;
    }
  CLV_ROWS_END
}

// reset_field as a buffer swap: after the device buffers of (x0, x1) have been exchanged, the cells OUTSIDE
// the update range (the halo rings, which reset_field_kernel_c.c:46-76 does not touch) are swapped back so
// that both arrays keep their own halos.  One launch for the four pairs.
struct RingPairs {
  double* a[4];
  double* b[4];
  int ext[4];  // 0 cell data (update range 1..nx x 1..ny), 1 vertex data (1..nx+1 x 1..ny+1)
};
__global__ void __launch_bounds__(256) ring_swap_kernel(RingPairs P, int nx, int ny, int pitch, unsigned long long* trace) {
  trace_min(trace, 0);
  pdl_wait();     // (programmatic dependent launch, common.cuh: the launch latency overlaps the previous kernel's tail)
  pdl_trigger();
  const int e = P.ext[blockIdx.y];
  const int W = nx + 4 + e, H = ny + 4 + e;        // Fortran extent (-1..nx+2+e) x (-1..ny+2+e)
  const int n_bt = 4 * W, n_lr = 4 * (H - 4);      // two full rows below + above, two columns left + right
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_bt + n_lr) {
    trace_max(trace, 1);
    return;
  }
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
  double* __restrict__ a = P.a[blockIdx.y];
  double* __restrict__ b = P.b[blockIdx.y];
  const double va = a[i], vb = b[i];
  a[i] = vb;
  b[i] = va;
  trace_max(trace, 1);
}

// copy of one array's update range (materialisation of a lazy copy, runtime.cu)
void launch_copy_range(const Grid& g, const double* src, double* dst, Kind kind) {
  const int e = (kind == VERTEX) ? 1 : 0;
  if (kind != CELL && kind != VERTEX) fatal("lazy copy of a non cell/vertex array");
  const Range r = make_range(1, g.nx + e, 1, g.ny + e);
  LaunchScope ls("lazy_copy");
  copy2_kernel<NR_COPY><<<grid_for(r, NR_COPY), dim3(BX, BY), 0, stream()>>>(r, g.pitch, src, dst, src, dst);
}

// ---- the calls on their own (host side): look up the device mirrors, launch ---------------------------
void run_ideal_gas(const Grid& g, double* density, double* energy, double* pressure, double* soundspeed) {
  const double* d = dev(g, density, CELL, IN);
  const double* e = dev(g, energy, CELL, IN);
  double* p = dev(g, pressure, CELL, OUT_FULL);
  double* ss = dev(g, soundspeed, CELL, OUT_FULL);
  const Range r = make_range(1, g.nx, 1, g.ny);
  LaunchScope ls("ideal_gas");
  ideal_gas_kernel<NR_IDEAL><<<grid_for(r, NR_IDEAL), dim3(BX, BY), 0, stream()>>>(r, g.pitch, d, e, p, ss);
}

void launch_soundspeed(const Grid& g, const double* density, const double* energy, double* soundspeed) {
  const Range r = make_range(1, g.nx, 1, g.ny);
  LaunchScope ls("soundspeed_lazy");
  soundspeed_kernel<<<grid_for(r, 1), dim3(BX, BY), 0, stream()>>>(r, g.pitch, density, energy, soundspeed);
}

void run_viscosity(const Grid& g, double* celldx, double* celldy, double* density0, double* pressure,
                   double* viscosity, double* xvel0, double* yvel0) {
  const double* cdx = dev(g, celldx, X1D_CELL, IN);
  const double* cdy = dev(g, celldy, Y1D_CELL, IN);
  const double* d0 = dev(g, density0, CELL, IN);
  const double* p = dev(g, pressure, CELL, IN);
  double* q = dev(g, viscosity, CELL, OUT_FULL);
  const double* xv = dev(g, xvel0, VERTEX, IN);
  const double* yv = dev(g, yvel0, VERTEX, IN);
  const Range r = make_range(1, g.nx, 1, g.ny);
  LaunchScope ls("viscosity");
  viscosity_kernel<NR_VISC><<<grid_for(r, NR_VISC), dim3(BX, BY), 0, stream()>>>(r, g.pitch, cdx, cdy, d0, p, q, xv, yv);
}

// the result lands in host_scalars()[0]; host_scalars()[7] then shows dt_result_seq()
double g_dt_seq = 0;
double dt_result_seq() { return g_dt_seq; }
void set_dt_result_seq(double s) { g_dt_seq = s; }
void run_calc_dt(const Grid& g, const DtParams& P, double* xarea, double* yarea, double* celldx, double* celldy,
                 double* volume, double* density0, double* viscosity, double* soundspeed, double* xvel0,
                 double* yvel0) {
  const double* xa = dev(g, xarea, XFACE, IN);
  const double* ya = dev(g, yarea, YFACE, IN);
  const double* cdx = dev(g, celldx, X1D_CELL, IN);
  const double* cdy = dev(g, celldy, Y1D_CELL, IN);
  const double* vol = dev(g, volume, CELL, IN);
  const double* d0 = dev(g, density0, CELL, IN);
  const double* q = dev(g, viscosity, CELL, IN);
  const double* ss = dev(g, soundspeed, CELL, IN);
  const double* xv = dev(g, xvel0, VERTEX, IN);
  const double* yv = dev(g, yvel0, VERTEX, IN);
  const Range r = make_range(1, g.nx, 1, g.ny);
  const dim3 grid = persistent_grid(r, NR_DT, 6);
  double* part = partials((size_t)grid.x * grid.y);
  const ReduceTail RT = next_reduce_tail(0);
  g_dt_seq = RT.seq;
  {
    LaunchScope ls("calc_dt");
    calc_dt_kernel<NR_DT><<<grid, dim3(BX, BY), 0, stream()>>>(r, g.pitch, P, xa, ya, cdx, cdy, vol, d0, q, ss, xv, yv,
                                                               part, ticket(), host_scalars(), RT);
  }
  note_fused_allreduce(0, 1, true, RT.all != nullptr);
}

void run_pdv(const Grid& g, bool predict, double dt, double* xarea, double* yarea, double* volume, double* density0,
             double* density1, double* energy0, double* energy1, double* pressure, double* viscosity,
             double* xvel0, double* xvel1, double* yvel0, double* yvel1) {
  const double* xa = dev(g, xarea, XFACE, IN);
  const double* ya = dev(g, yarea, YFACE, IN);
  const double* vol = dev(g, volume, CELL, IN);
  const double* d0 = dev(g, density0, CELL, IN);
  double* d1 = dev(g, density1, CELL, OUT_FULL);
  const double* e0 = dev(g, energy0, CELL, IN);
  double* e1 = dev(g, energy1, CELL, OUT_FULL);
  const double* p = dev(g, pressure, CELL, IN);
  const double* q = dev(g, viscosity, CELL, IN);
  const double* x0 = dev(g, xvel0, VERTEX, IN);
  const double* y0 = dev(g, yvel0, VERTEX, IN);
  const Range r = make_range(1, g.nx, 1, g.ny);
  const dim3 grid = grid_for(r, NR_PDV);
  if (predict) {
    LaunchScope ls("pdv_predict");
    pdv_kernel<true, NR_PDV><<<grid, dim3(BX, BY), 0, stream()>>>(r, g.pitch, dt, xa, ya, vol, d0, d1, e0, e1, p, q, x0,
                                                                  x0, y0, y0);
  } else {
    const double* x1 = dev(g, xvel1, VERTEX, IN);
    const double* y1 = dev(g, yvel1, VERTEX, IN);
    LaunchScope ls("pdv_correct");
    pdv_kernel<false, NR_PDV><<<grid, dim3(BX, BY), 0, stream()>>>(r, g.pitch, dt, xa, ya, vol, d0, d1, e0, e1, p, q,
                                                                   x0, x1, y0, y1);
  }
}

void run_revert(const Grid& g, double* density0, double* density1, double* energy0, double* energy1) {
  if (is_resident()) {  // recorded, not copied: the corrector overwrites both before anything reads them
    lazy_copy(g, density1, density0, CELL);
    lazy_copy(g, energy1, energy0, CELL);
    return;
  }
  const double* d0 = dev(g, density0, CELL, IN);
  double* d1 = dev(g, density1, CELL, OUT);
  const double* e0 = dev(g, energy0, CELL, IN);
  double* e1 = dev(g, energy1, CELL, OUT);
  const Range r = make_range(1, g.nx, 1, g.ny);
  LaunchScope ls("revert");
  copy2_kernel<NR_COPY><<<grid_for(r, NR_COPY), dim3(BX, BY), 0, stream()>>>(r, g.pitch, d0, d1, e0, e1);
}

void run_reset_field(const Grid& g, double* density0, double* density1, double* energy0, double* energy1,
                     double* xvel0, double* xvel1, double* yvel0, double* yvel1) {
  if (is_resident()) {
    // x0 := x1 on the update range.  Swap the two device buffers, swap the halo rings back, and record
    // "x1 := x0" as a lazy copy (the next PdV / accelerate overwrites x1 before anything reads it).
    double* h0[4] = {density0, energy0, xvel0, yvel0};
    double* h1[4] = {density1, energy1, xvel1, yvel1};
    const Kind kinds[4] = {CELL, CELL, VERTEX, VERTEX};
    RingPairs P;
    for (int i = 0; i < 4; ++i) {
      dev(g, h1[i], kinds[i], IN);     // the source really holds its data
      dev(g, h0[i], kinds[i], INOUT);  // nothing pending on or from the destination
      swap_buffers(h0[i], h1[i]);
      P.a[i] = dev(g, h0[i], kinds[i], INOUT);
      P.b[i] = dev(g, h1[i], kinds[i], INOUT);
      P.ext[i] = kinds[i] == VERTEX ? 1 : 0;
    }
    {
      const int ring = 4 * (g.nx + 5) + 4 * (g.ny + 1);
      LaunchScope ls("reset_field_swap");
      launch_pdl(ring_swap_kernel, dim3((unsigned)((ring + 255) / 256), 4), dim3(256), 0, stream(), P, g.nx, g.ny, g.pitch, ls.trace);
    }
    note_ring_swap_launch();
    for (int i = 0; i < 4; ++i) lazy_copy(g, h1[i], h0[i], kinds[i]);
    return;
  }
  double* d0 = dev(g, density0, CELL, OUT);
  const double* d1 = dev(g, density1, CELL, IN);
  double* e0 = dev(g, energy0, CELL, OUT);
  const double* e1 = dev(g, energy1, CELL, IN);
  double* x0 = dev(g, xvel0, VERTEX, OUT);
  const double* x1 = dev(g, xvel1, VERTEX, IN);
  double* y0 = dev(g, yvel0, VERTEX, OUT);
  const double* y1 = dev(g, yvel1, VERTEX, IN);
  const Range r = make_range(1, g.nx + 1, 1, g.ny + 1);
  LaunchScope ls("reset_field");
  reset_field_kernel<NR_RESET><<<grid_for(r, NR_RESET), dim3(BX, BY), 0, stream()>>>(r, g.pitch, g.nx, g.ny, d0, d1, e0,
                                                                                  e1, x0, x1, y0, y1);
}

void run_accelerate(const Grid& g, double dt, double* xarea, double* yarea, double* volume, double* density0,
                    double* pressure, double* viscosity, double* xvel0, double* yvel0, double* xvel1,
                    double* yvel1) {
  const double* xa = dev(g, xarea, XFACE, IN);
  const double* ya = dev(g, yarea, YFACE, IN);
  const double* vol = dev(g, volume, CELL, IN);
  const double* d0 = dev(g, density0, CELL, IN);
  const double* p = dev(g, pressure, CELL, IN);
  const double* q = dev(g, viscosity, CELL, IN);
  const double* x0 = dev(g, xvel0, VERTEX, IN);
  const double* y0 = dev(g, yvel0, VERTEX, IN);
  double* x1 = dev(g, xvel1, VERTEX, OUT_FULL);
  double* y1 = dev(g, yvel1, VERTEX, OUT_FULL);
  const Range r = make_range(1, g.nx + 1, 1, g.ny + 1);
  LaunchScope ls("accelerate");
  accelerate_kernel<NR_ACC><<<grid_for(r, NR_ACC), dim3(BX, BY), 0, stream()>>>(r, g.pitch, dt, xa, ya, vol, d0, p, q,
                                                                              x0, y0, x1, y1);
}

void run_flux_calc(const Grid& g, double dt, double* xarea, double* yarea, double* xvel0, double* yvel0,
                   double* xvel1, double* yvel1, double* vol_flux_x, double* vol_flux_y) {
  const double* xa = dev(g, xarea, XFACE, IN);
  const double* ya = dev(g, yarea, YFACE, IN);
  const double* x0 = dev(g, xvel0, VERTEX, IN);
  const double* y0 = dev(g, yvel0, VERTEX, IN);
  const double* x1 = dev(g, xvel1, VERTEX, IN);
  const double* y1 = dev(g, yvel1, VERTEX, IN);
  double* fx = dev(g, vol_flux_x, XFACE, OUT_FULL);
  double* fy = dev(g, vol_flux_y, YFACE, OUT_FULL);
  const Range r = make_range(1, g.nx + 1, 1, g.ny + 1);
  LaunchScope ls("flux_calc");
  flux_calc_kernel<NR_FLUX><<<grid_for(r, NR_FLUX), dim3(BX, BY), 0, stream()>>>(r, g.pitch, g.nx, g.ny, dt, xa, ya, x0,
                                                                               y0, x1, y1, fx, fy);
}

}  // namespace clv

using namespace clv;

extern "C" {

// Op::a = {density, energy, pressure, soundspeed}
void ideal_gas_kernel_c_(int* xmin, int* xmax, int* ymin, int* ymax, double* density, double* energy,
                         double* pressure, double* soundspeed) {
  Op op;
  op.kind = OP_IDEAL_GAS;
  const Grid g = op.g = grid_of_noflush(xmin, xmax, ymin, ymax);
  double* a[] = {density, energy, pressure, soundspeed};
  for (double* p : a) op.a[op.na++] = p;
  op.reads({density, energy});
  op.overwrites({pressure, soundspeed});
  op.run = [=] { run_ideal_gas(g, density, energy, pressure, soundspeed); };
  submit(std::move(op));
}

// Op::a = {celldx, celldy, density0, pressure, viscosity, xvel0, yvel0}
void viscosity_kernel_c_(int* xmin, int* xmax, int* ymin, int* ymax, double* celldx, double* celldy,
                         double* density0, double* pressure, double* viscosity, double* xvel0,
                         double* yvel0) {
  Op op;
  op.kind = OP_VISCOSITY;
  const Grid g = op.g = grid_of_noflush(xmin, xmax, ymin, ymax);
  double* a[] = {celldx, celldy, density0, pressure, viscosity, xvel0, yvel0};
  for (double* p : a) op.a[op.na++] = p;
  op.reads({celldx, celldy, density0, pressure, xvel0, yvel0});
  op.overwrites({viscosity});
  op.run = [=] { run_viscosity(g, celldx, celldy, density0, pressure, viscosity, xvel0, yvel0); };
  submit(std::move(op));
}

// Op::a = {xarea, yarea, celldx, celldy, volume, density0, viscosity, soundspeed, xvel0, yvel0}; sv = DtParams
void calc_dt_kernel_c_(int* xmin, int* xmax, int* ymin, int* ymax, double* g_small, double* g_big,
                       double* dtmin, double* dtc_safe, double* dtu_safe, double* dtv_safe,
                       double* dtdiv_safe, double* xarea, double* yarea, double* cellx, double* celly,
                       double* celldx, double* celldy, double* volume, double* density0,
                       double* energy0, double* pressure, double* viscosity, double* soundspeed,
                       double* xvel0, double* yvel0, double* dt_min, double* dt_min_val,
                       int* dtl_control, double* xl_pos, double* yl_pos, int* jldt, int* kldt,
                       int* small) {
  (void)cellx; (void)celly; (void)energy0; (void)pressure; (void)dt_min; (void)xl_pos; (void)yl_pos;
  Op op;
  op.kind = OP_CALC_DT;
  const Grid g = op.g = grid_of_noflush(xmin, xmax, ymin, ymax);
  double* a[] = {xarea, yarea, celldx, celldy, volume, density0, viscosity, soundspeed, xvel0, yvel0};
  for (double* p : a) op.a[op.na++] = p;
  const DtParams P{*g_small, *g_big, *dtc_safe, *dtu_safe, *dtv_safe, *dtdiv_safe};
  op.sv[0] = P.g_small; op.sv[1] = P.g_big; op.sv[2] = P.dtc_safe; op.sv[3] = P.dtu_safe; op.sv[4] = P.dtv_safe;
  op.sv[5] = P.dtdiv_safe;
  op.reads({xarea, yarea, celldx, celldy, volume, density0, viscosity, soundspeed, xvel0, yvel0});
  op.run = [=] { run_calc_dt(g, P, xarea, yarea, celldx, celldy, volume, density0, viscosity, soundspeed, xvel0, yvel0); };
  submit(std::move(op));
  flush_deferred();
  // the one unavoidable host-visible result per step: the host spins on the sequence number the reduction tail
  // writes to pinned memory after the value (no stream synchronisation: launches that follow calc_dt in the
  // stream, e.g. the viscosity halo exchange, keep running).  With several ranks the kernel has already folded the
  // minimum across ranks (host_scalars()[0]); calc_dt's own contract is the LOCAL minimum ([32]); clover_b200_min_
  // hands out the global one.
  wait_scalars(0, dt_result_seq());
  const double v = host_scalars()[32];
  *dt_min_val = v;
  *dtl_control = 1;  // calc_dt_kernel_c.c:159-163
  *jldt = 1;
  *kldt = 1;
  if (v < *dtmin) {
    if (small) *small = 1;
    printf("Timestep information:\ntimestep : %f (below dtmin)\n", v);
  }
}

// Op::a = {xarea, yarea, volume, density0, density1, energy0, energy1, pressure, viscosity, xvel0, xvel1, yvel0,
//          yvel1}; sv[0] = dt
void pdv_kernel_c_(int* prdct, int* xmin, int* xmax, int* ymin, int* ymax, double* dt, double* xarea,
                   double* yarea, double* volume, double* density0, double* density1, double* energy0,
                   double* energy1, double* pressure, double* viscosity, double* xvel0, double* xvel1,
                   double* yvel0, double* yvel1, double* volume_change) {
  (void)volume_change;
  Op op;
  const bool predict = (*prdct == 0);
  op.kind = predict ? OP_PDV_PREDICT : OP_PDV_CORRECT;
  const Grid g = op.g = grid_of_noflush(xmin, xmax, ymin, ymax);
  double* a[] = {xarea, yarea, volume, density0, density1, energy0, energy1, pressure, viscosity, xvel0, xvel1,
                 yvel0, yvel1};
  for (double* p : a) op.a[op.na++] = p;
  const double dtv = op.sv[0] = *dt;
  op.reads({xarea, yarea, volume, density0, energy0, pressure, viscosity, xvel0, yvel0});
  if (!predict) op.reads({xvel1, yvel1});
  op.overwrites({density1, energy1});
  op.run = [=] {
    run_pdv(g, predict, dtv, xarea, yarea, volume, density0, density1, energy0, energy1, pressure, viscosity, xvel0,
            xvel1, yvel0, yvel1);
  };
  submit(std::move(op));
}

// Op::a = {density0, density1, energy0, energy1}
void revert_kernel_c_(int* xmin, int* xmax, int* ymin, int* ymax, double* density0, double* density1,
                      double* energy0, double* energy1) {
  Op op;
  op.kind = OP_REVERT;
  const Grid g = op.g = grid_of_noflush(xmin, xmax, ymin, ymax);
  double* a[] = {density0, density1, energy0, energy1};
  for (double* p : a) op.a[op.na++] = p;
  op.reads({density0, energy0});
  op.overwrites({density1, energy1});
  op.run = [=] { run_revert(g, density0, density1, energy0, energy1); };
  submit(std::move(op));
}

// Op::a = {density0, density1, energy0, energy1, xvel0, xvel1, yvel0, yvel1}
void reset_field_kernel_c_(int* xmin, int* xmax, int* ymin, int* ymax, double* density0,
                           double* density1, double* energy0, double* energy1, double* xvel0,
                           double* xvel1, double* yvel0, double* yvel1) {
  Op op;
  op.kind = OP_RESET_FIELD;
  const Grid g = op.g = grid_of_noflush(xmin, xmax, ymin, ymax);
  double* a[] = {density0, density1, energy0, energy1, xvel0, xvel1, yvel0, yvel1};
  for (double* p : a) op.a[op.na++] = p;
  op.reads({density1, energy1, xvel1, yvel1});
  op.overwrites({density0, energy0, xvel0, yvel0});
  op.run = [=] { run_reset_field(g, density0, density1, energy0, energy1, xvel0, xvel1, yvel0, yvel1); };
  submit(std::move(op));
}

// Op::a = {xarea, yarea, volume, density0, pressure, viscosity, xvel0, yvel0, xvel1, yvel1}; sv[0] = dt
void accelerate_kernel_c_(int* xmin, int* xmax, int* ymin, int* ymax, double* dt, double* xarea,
                          double* yarea, double* volume, double* density0, double* pressure,
                          double* viscosity, double* xvel0, double* yvel0, double* xvel1,
                          double* yvel1) {
  Op op;
  op.kind = OP_ACCELERATE;
  const Grid g = op.g = grid_of_noflush(xmin, xmax, ymin, ymax);
  double* a[] = {xarea, yarea, volume, density0, pressure, viscosity, xvel0, yvel0, xvel1, yvel1};
  for (double* p : a) op.a[op.na++] = p;
  const double dtv = op.sv[0] = *dt;
  op.reads({xarea, yarea, volume, density0, pressure, viscosity, xvel0, yvel0});
  op.overwrites({xvel1, yvel1});
  op.run = [=] { run_accelerate(g, dtv, xarea, yarea, volume, density0, pressure, viscosity, xvel0, yvel0, xvel1, yvel1); };
  submit(std::move(op));
}

// Op::a = {xarea, yarea, xvel0, yvel0, xvel1, yvel1, vol_flux_x, vol_flux_y}; sv[0] = dt
void flux_calc_kernel_c_(int* xmin, int* xmax, int* ymin, int* ymax, double* dt, double* xarea,
                         double* yarea, double* xvel0, double* yvel0, double* xvel1, double* yvel1,
                         double* vol_flux_x, double* vol_flux_y) {
  Op op;
  op.kind = OP_FLUX_CALC;
  const Grid g = op.g = grid_of_noflush(xmin, xmax, ymin, ymax);
  double* a[] = {xarea, yarea, xvel0, yvel0, xvel1, yvel1, vol_flux_x, vol_flux_y};
  for (double* p : a) op.a[op.na++] = p;
  const double dtv = op.sv[0] = *dt;
  op.reads({xarea, yarea, xvel0, yvel0, xvel1, yvel1});
  op.writes({vol_flux_x, vol_flux_y});
  op.run = [=] { run_flux_calc(g, dtv, xarea, yarea, xvel0, yvel0, xvel1, yvel1, vol_flux_x, vol_flux_y); };
  submit(std::move(op));
}

void field_summary_kernel_c_(int* xmin, int* xmax, int* ymin, int* ymax, double* volume,
                             double* density0, double* energy0, double* pressure, double* xvel0,
                             double* yvel0, double* vol, double* mass, double* ie, double* ke,
                             double* press) {
  const Grid g = grid_of(xmin, xmax, ymin, ymax);  // drains the queue
  const double* v = dev(g, volume, CELL, IN);
  const double* d0 = dev(g, density0, CELL, IN);
  const double* e0 = dev(g, energy0, CELL, IN);
  const double* p = dev(g, pressure, CELL, IN);
  const double* x0 = dev(g, xvel0, VERTEX, IN);
  const double* y0 = dev(g, yvel0, VERTEX, IN);
  const Range r = make_range(1, g.nx, 1, g.ny);
  const dim3 grid = persistent_grid(r, NR_SUM, 6);
  double* part = partials((size_t)grid.x * grid.y * 5);
  double* out = host_scalars() + 8;
  const ReduceTail RT = next_reduce_tail(8);
  {
    LaunchScope ls("field_summary");
    field_summary_kernel<NR_SUM><<<grid, dim3(BX, BY), 0, stream()>>>(r, g.pitch, v, d0, e0, p, x0, y0, part,
                                                                      ticket() + 1, out, RT);
  }
  note_fused_allreduce(8, 5, false, RT.all != nullptr);
  wait_scalars(8, RT.seq);
  // this rank's sums (the kernel's contract); the sums over all ranks are already in out[0..4] for clover_b200_sum_
  *vol = out[32 + 0];
  *mass = out[32 + 1];
  *ie = out[32 + 2];
  *ke = out[32 + 3];
  *press = out[32 + 4];
  finish();
}

}  // extern "C"
