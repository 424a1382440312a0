// lagrange.cuh -- device-side building blocks shared by lagrange.cu (one kernel per reference call) and
// fuse.cu (several reference calls in one kernel): tile/row mapping macros, the per-cell arithmetic of
// ideal_gas, viscosity and calc_dt in the reference's evaluation order, and the block-reduction tail.
#pragma once
#include "common.cuh"

namespace clv {

constexpr int BX = 32, BY = 8;
// rows per thread, per kernel (tuned on B200, see profiles/)
constexpr int NR_IDEAL = 4, NR_VISC = 1, NR_DT = 1, NR_PDV = 2, NR_COPY = 4, NR_RESET = 1, NR_ACC = 1, NR_FLUX = 2,
              NR_SUM = 2;

struct Range {
  int j0, j1, k0, k1;  // inclusive
  int jbase;           // first column handled by block x = 0 (16-double aligned)
};

static inline Range make_range(int j0, int j1, int k0, int k1) {
  Range r{j0, j1, k0, k1, 0};
  r.jbase = ((j0 + XOFF) & ~15) - XOFF;
  return r;
}
static inline dim3 grid_for(const Range& r, int nr) {
  return dim3((unsigned)((r.j1 - r.jbase + 1 + BX - 1) / BX),
              (unsigned)((r.k1 - r.k0 + 1 + BY * nr - 1) / (BY * nr)), 1);
}
// Opens the unrolled row loop: defines j, k (clamped into the range, always safe to load from) and
// `active` (this thread really owns (j,k): predicate for stores / reductions).
#define CLV_ROWS_BEGIN(r, NR)                                                          \
  const int j_raw_ = (r).jbase + (int)(blockIdx.x * BX + threadIdx.x);                 \
  const bool j_ok_ = (j_raw_ >= (r).j0) && (j_raw_ <= (r).j1);                         \
  const int j = j_raw_ < (r).j0 ? (r).j0 : (j_raw_ > (r).j1 ? (r).j1 : j_raw_);        \
  _Pragma("unroll") for (int rr_ = 0; rr_ < (NR); ++rr_) {                             \
    const int k_raw_ = (r).k0 + (int)((blockIdx.y * (NR) + rr_) * BY + threadIdx.y);   \
    const bool active = j_ok_ && (k_raw_ <= (r).k1);                                   \
    const int k = k_raw_ <= (r).k1 ? k_raw_ : (r).k1;
#define CLV_ROWS_END }
// Persistent variant for the reduction kernels: a 1-D grid of a few CTAs per SM walks the same 32x(8*NR)
// tiles in a grid-stride loop, so that the block-level reduction tail (fence + ticket atomic) is paid once
// per CTA instead of once per tile (it held every warp of a 256-cell block hostage for ~2k cycles).
#define CLV_PTILES_BEGIN(r, NR)                                                        \
  const unsigned tiles_x_ = (unsigned)(((r).j1 - (r).jbase + BX) / BX);                \
  const unsigned tiles_y_ = (unsigned)(((r).k1 - (r).k0 + BY * (NR)) / (BY * (NR)));   \
  for (unsigned tile_ = blockIdx.x; tile_ < tiles_x_ * tiles_y_; tile_ += gridDim.x) { \
    const unsigned bx_ = tile_ % tiles_x_, by_ = tile_ / tiles_x_;                     \
    const int j_raw_ = (r).jbase + (int)(bx_ * BX + threadIdx.x);                      \
    const bool j_ok_ = (j_raw_ >= (r).j0) && (j_raw_ <= (r).j1);                       \
    const int j = j_raw_ < (r).j0 ? (r).j0 : (j_raw_ > (r).j1 ? (r).j1 : j_raw_);      \
    _Pragma("unroll") for (int rr_ = 0; rr_ < (NR); ++rr_) {                           \
      const int k_raw_ = (r).k0 + (int)((by_ * (NR) + rr_) * BY + threadIdx.y);        \
      const bool active = j_ok_ && (k_raw_ <= (r).k1);                                 \
      const int k = k_raw_ <= (r).k1 ? k_raw_ : (r).k1;
#define CLV_PTILES_END }}
static inline dim3 persistent_grid(const Range& r, int nr, int ctas_per_sm) {
  const dim3 g = grid_for(r, nr);
  const unsigned tiles = g.x * g.y, cap = 148u * (unsigned)ctas_per_sm;
  return dim3(tiles < cap ? tiles : cap, 1, 1);
}

// ------------------------------------------------------------------------------------------------
// ideal_gas_kernel_c.c:48-59.  4 passes (2 reads, 2 writes) = 32 B/cell.
template <bool SAFE>
__device__ __forceinline__ void ideal_gas_cell(double rho, double e, double& p, double& ss, bool& bad) {
  const double v = Math<SAFE>::rcp(rho, bad);
  p = (1.4 - 1.0) * rho * e;
  const double pe = (1.4 - 1.0) * rho;
  const double pv = -rho * p;
  const double ss2 = v * v * (p * pe - pv);
  ss = Math<SAFE>::sqrt(ss2, bad);
}

// ------------------------------------------------------------------------------------------------
// viscosity_kernel_c.c:53-104.  5 passes = 40 B/cell.
struct ViscIn {
  double u00, u10, u01, u11, v00, v10, v01, v11, dx, dy, dx1, dy1, pl, pr, pb, pt, rho;
};
template <bool SAFE>
__device__ __forceinline__ double viscosity_cell(const ViscIn& I, bool& bad) {
  typedef Math<SAFE> M;
  const double ugrad = (I.u10 + I.u11) - (I.u00 + I.u01);
  const double vgrad = (I.v01 + I.v11) - (I.v00 + I.v10);
  const double div = I.dx * ugrad + I.dy * vgrad;
  // viscosity_kernel_c.c:88: `if (limiter>0.0 || div>=0.0) viscosity = 0`.  The limiter (7 divisions)
  // only decides anything for a compressing cell, so it is evaluated only there; everywhere else --
  // the whole quiescent part of the mesh -- the answer is 0 whatever the limiter is.
  if (div >= 0.0) return 0.0;
  const double strain2 = M::div(0.5 * (I.u01 + I.u11 - I.u00 - I.u10), I.dy, bad) +
                         M::div(0.5 * (I.v10 + I.v11 - I.v00 - I.v01), I.dx, bad);
  double pgradx = M::div(I.pr - I.pl, I.dx + I.dx1, bad);
  double pgrady = M::div(I.pt - I.pb, I.dy + I.dy1, bad);
  const double pgradx2 = pgradx * pgradx, pgrady2 = pgrady * pgrady;
  const double limiter = M::div(M::div(0.5 * ugrad, I.dx, bad) * pgradx2 + M::div(0.5 * vgrad, I.dy, bad) * pgrady2 +
                                    strain2 * pgradx * pgrady,
                                dmax(pgradx2 + pgrady2, 1.0e-16), bad);
  double q = 0.0;
  if (!(limiter > 0.0)) {  // generic operators in this minority branch
    const double ax = dmax(1.0e-16, fabs(pgradx)), ay = dmax(1.0e-16, fabs(pgrady));
    pgradx = (pgradx < 0.0) ? -ax : ax;
    pgrady = (pgrady < 0.0) ? -ay : ay;
    const double pgrad = sqrt(pgradx * pgradx + pgrady * pgrady);
    const double xgrad = fabs(I.dx * pgrad / pgradx);
    const double ygrad = fabs(I.dy * pgrad / pgrady);
    const double grad = dmin(xgrad, ygrad);
    const double grad2 = grad * grad;
    q = 2.0 * I.rho * grad2 * limiter * limiter;
  }
  return q;
}


// ------------------------------------------------------------------------------------------------
// Block-level reduction tails shared by calc_dt and field_summary: every block publishes its
// partial(s), the last block to arrive (ticket) folds them in a fixed order.  With several ranks
// (ReduceTail.all != nullptr) the same block then folds ACROSS ranks over peer memory -- thread r drops this
// rank's values + sequence number into its mailbox in rank r's block and waits for rank r's (bounded spin), thread 0
// folds in rank order, so every rank holds bit-identical results -- which replaces clover_min / clover_sum
// (clover.f90:3621-3657) without a launch or a host round trip of their own.  Results go to pinned host memory:
//   out[0..N)  the (global) result    out[32..32+N)  this rank's local result    out[7]  sequence number, written
// last (the host spins on it instead of synchronising the stream; runtime.cu: wait_scalars).
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded system-scope spin: returns once *flag >= want; after timeout_ns writes an error record
// {code, rank, who, want, seen} to err (pinned host memory, printed by the host's CUDA error path) and traps.
__device__ __forceinline__ void spin_until_ge(const unsigned long long* flag, unsigned long long want,
                                              unsigned long long timeout_ns, double* err, int code, int rank, int who) {
  unsigned long long v, t0 = 0;
  unsigned int polls = 0;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= want) return;
    if ((++polls & 1023u) == 0) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > timeout_ns) break;
    }
  }
  if (err) {
    err[1] = (double)rank; err[2] = (double)who; err[3] = (double)want; err[4] = (double)v;
    __threadfence_system();
    err[0] = (double)code;
    __threadfence_system();
  }
  __trap();
}

template <int N, bool IS_MIN>
__device__ __forceinline__ void block_reduce_publish(double (&v)[N], double* __restrict__ partials,
                                                     unsigned int* ticket, double* __restrict__ out,
                                                     double identity, const ReduceTail& RT) {
  __shared__ double sm[N][BX * BY / 32];
  __shared__ double xr[RT_MAX_RANKS][N];
  __shared__ bool last;
  const int tid = threadIdx.y * BX + threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const unsigned nblocks = gridDim.x * gridDim.y;
  const unsigned bid = blockIdx.y * gridDim.x + blockIdx.x;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double w = IS_MIN ? warp_min(v[i]) : warp_sum(v[i]);
    if (lane == 0) sm[i][warp] = w;
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double a = sm[i][0];
      for (int w = 1; w < BX * BY / 32; ++w) a = IS_MIN ? ((sm[i][w] < a) ? sm[i][w] : a) : a + sm[i][w];
      partials[(size_t)i * nblocks + bid] = a;
    }
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    last = (t == nblocks - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double a = identity;
    for (unsigned b = tid; b < nblocks; b += BX * BY) {
      const double p = __ldcg(&partials[(size_t)i * nblocks + b]);
      a = IS_MIN ? ((p < a) ? p : a) : a + p;
    }
    a = IS_MIN ? warp_min(a) : warp_sum(a);
    if (lane == 0) sm[i][warp] = a;
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double a = sm[i][0];
      for (int w = 1; w < BX * BY / 32; ++w) a = IS_MIN ? ((sm[i][w] < a) ? sm[i][w] : a) : a + sm[i][w];
      sm[i][0] = a;  // this rank's result
      out[32 + i] = a;
    }
    *ticket = 0;
  }
  __syncthreads();
  if (RT.all != nullptr) {
    // ---- across ranks, over peer memory (mailbox [parity][sender] of RT_SLOT bytes: 8 values + sequence number)
    const size_t box = RT_OFF + (size_t)(RT.ar_seq & 1) * RT_MAX_RANKS * RT_SLOT;
    if (tid < RT.nranks) {
      double* dst = reinterpret_cast<double*>(RT.all[tid] + box + (size_t)RT.rank * RT_SLOT);
#pragma unroll
      for (int i = 0; i < N; ++i) dst[i] = sm[i][0];
      __threadfence_system();
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst + 8), "l"(RT.ar_seq) : "memory");
      const double* src = reinterpret_cast<const double*>(RT.all[RT.rank] + box + (size_t)tid * RT_SLOT);
      spin_until_ge(reinterpret_cast<const unsigned long long*>(src + 8), RT.ar_seq, RT.timeout_ns, RT.err, 3, RT.rank, tid);
#pragma unroll
      for (int i = 0; i < N; ++i) xr[tid][i] = __ldcg(src + i);
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double a = xr[0][i];
        for (int q = 1; q < RT.nranks; ++q) a = IS_MIN ? ((xr[q][i] < a) ? xr[q][i] : a) : a + xr[q][i];
        sm[i][0] = a;
      }
    }
  }
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = sm[i][0];
    __threadfence_system();
    out[7] = RT.seq;
    __threadfence_system();
  }
}

// calc_dt_kernel_c.c:96-144.  8 passes read = 64 B/cell; no per-cell dt_min array is written.
struct DtParams {
  double g_small, g_big, dtc_safe, dtu_safe, dtv_safe, dtdiv_safe;
};
struct DtIn {
  double dsx, dsy, vol, ssp, visc, rho, u00, u10, u01, u11, v00, v10, v01, v11, xa0, xa1, ya0, ya1;
};
// calc_dt_kernel_c.c:99-133, one cell.  The sqrt and the four divisions are independent chains.
template <bool SAFE>
__device__ __forceinline__ double calc_dt_cell(const DtIn& I, const DtParams& P, bool& bad) {
  typedef Math<SAFE> M;
  double cc = I.ssp * I.ssp;
  cc = cc + M::div(2.0 * I.visc, I.rho, bad);
  cc = dmax(M::sqrt(cc, bad), P.g_small);
  const double dtct = M::div(P.dtc_safe * dmin(I.dsx, I.dsy), cc, bad);
  double div = 0.0;
  double dv1 = (I.u00 + I.u01) * I.xa0;
  double dv2 = (I.u10 + I.u11) * I.xa1;
  div = div + dv2 - dv1;
  const double dtut = M::div(P.dtu_safe * 2.0 * I.vol, dmax(fabs(dv1), dmax(fabs(dv2), P.g_small * I.vol)), bad);
  dv1 = (I.v00 + I.v10) * I.ya0;
  dv2 = (I.v01 + I.v11) * I.ya1;
  div = div + dv2 - dv1;
  const double dtvt = M::div(P.dtv_safe * 2.0 * I.vol, dmax(fabs(dv1), dmax(fabs(dv2), P.g_small * I.vol)), bad);
  div = M::div(div, 2.0 * I.vol, bad);
  // the divergence limit applies to compressing cells only: generic operator in that minority branch
  const double dtdivt = (div < -P.g_small) ? P.dtdiv_safe * (-1.0 / div) : P.g_big;
  return dmin(dtct, dmin(dtut, dmin(dtvt, dtdivt)));
}

// The same minimum as calc_dt_cell, evaluated with fewer divisions (the fused timestep launch is issue-bound, not
// bandwidth-bound: profiles/).  Bit-identical by construction:
//   * 2q/rho is +-0 when q == 0 (rho is normal: ideal_gas's 1/rho has already flagged anything else), and cc + (+-0)
//     == cc for cc = c*c >= +0: the division is skipped when no lane of the warp has a non-zero viscosity;
//   * dtu = numu/denu only matters if it is the minimum.  If numu > dtct*denu*(1+1e-6) then the correctly rounded
//     quotient is >= dtct (two roundings of 2^-53 against a margin of 1e-6), so min(dtct, dtu, ..) == min(dtct, ..):
//     the division is skipped (dtu := g_big; the result never exceeds g_big because dtdiv <= g_big) unless some lane
//     of the warp cannot rule it out.  Same for dtv.  NaNs fail the test and take the division.
//   * div/(2 vol) >= 0 whenever div >= 0 (vol > 0), i.e. "not compressing": dtdiv = g_big without dividing.
// Lanes that do not need a division but sit in a warp that does simply compute it (same value as before).
template <bool SAFE>
__device__ __forceinline__ double calc_dt_cell_lean(const DtIn& I, const DtParams& P, bool& bad, unsigned mask) {
  typedef Math<SAFE> M;
  // (the generic re-run -- SAFE, only the lanes whose fast path flagged an operand -- does not vote: it divides)
  auto any = [&](bool pred) { return SAFE ? true : (__any_sync(mask, pred) != 0); };
  double cc = I.ssp * I.ssp;
  if (any(I.visc != 0.0)) cc = cc + M::div(2.0 * I.visc, I.rho, bad);
  cc = dmax(M::sqrt(cc, bad), P.g_small);
  const double dtct = M::div(P.dtc_safe * dmin(I.dsx, I.dsy), cc, bad);
  double div = 0.0;
  double dv1 = (I.u00 + I.u01) * I.xa0;
  double dv2 = (I.u10 + I.u11) * I.xa1;
  div = div + dv2 - dv1;
  const double numu = P.dtu_safe * 2.0 * I.vol;
  const double denu = dmax(fabs(dv1), dmax(fabs(dv2), P.g_small * I.vol));
  dv1 = (I.v00 + I.v10) * I.ya0;
  dv2 = (I.v01 + I.v11) * I.ya1;
  div = div + dv2 - dv1;
  const double numv = P.dtv_safe * 2.0 * I.vol;
  const double denv = dmax(fabs(dv1), dmax(fabs(dv2), P.g_small * I.vol));
  const double bu = dtct * denu, bv = dtct * denv;
  const bool need_u = !(bu >= 1.0e-280 && numu > bu * 1.000001);
  const bool need_v = !(bv >= 1.0e-280 && numv > bv * 1.000001);
  double dtut = P.g_big, dtvt = P.g_big, dtdivt = P.g_big;
  if (any(need_u)) dtut = M::div(numu, denu, bad);
  if (any(need_v)) dtvt = M::div(numv, denv, bad);
  if (any(!(div >= 0.0))) {
    const double dq = M::div(div, 2.0 * I.vol, bad);
    // the divergence limit applies to compressing cells only: generic operator in that minority branch
    dtdivt = (dq < -P.g_small) ? P.dtdiv_safe * (-1.0 / dq) : P.g_big;
  }
  return dmin(dtct, dmin(dtut, dmin(dtvt, dtdivt)));
}

}  // namespace clv
