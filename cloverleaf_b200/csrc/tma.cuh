// tma.cuh -- TMA tile staging for the streaming stencil kernels (sm_100a).
//
// Why: the fp64 kernels hold 50-130 registers per thread, so only 16-32 warps per SM are resident and the bytes
// they keep in flight with ordinary loads (~30 KB/SM) cannot cover HBM latency (profiles/r01b, r01e: DRAM 35-60 %,
// dominant stall long_scoreboard).  Here a persistent CTA walks over tiles of the chunk; ONE thread asks the TMA
// unit (cp.async.bulk.tensor.2d, SASS UTMALDG) to copy the stencil neighbourhood of a LATER tile -- a
// (TW+halo) x (TH+halo) box of every input field -- into a ring of shared-memory stages while all threads compute
// the current tile out of shared memory.  Bytes in flight are bounded by shared memory (up to ~150 KB/SM), cost no
// registers and no issue slots, and the address arithmetic / clamping of the halo loads disappears: boxes that
// stick out of the allocation are zero-filled by the hardware.
//
// Every field uses the same pitched layout (common.cuh), so one tensor-map shape serves all of them:
//     dim0 = pitch doubles (unit stride), dim1 = ny+6 rows, row stride = pitch*8 B (a multiple of 128 B);
//     element (j,k) sits at coordinates (j + XOFF, k + 1).
// Measured on B200: the first byte of a box must be 16-byte aligned (an odd dim-0 coordinate of an fp64 tensor
// raises cudaErrorIllegalInstruction), so boxes start at an ODD j (XOFF is odd): tiles start at j0 = 1 + n*TW
// with TW even, and a box that needs column j0-1 starts at j0-2.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace clv {

// Host: tensor map of `dev_ptr` (a field in the pitched layout of grid g) with a box_w x box_h box; cached.
const CUtensorMap* tensor_map_for(const Grid& g, const double* dev_ptr, int box_w, int box_h);

// Host: device table of the ntx*nty tile coordinates in the order the persistent CTAs' ticket queue hands them out;
// cached per shape (runtime.cu; the order itself is plain C++ in tile_order.h).  Chunks up to ~4096 cells wide are
// listed row by row: the CTAs that run at the same time then stream long contiguous row segments and the halo rows
// shared with the next tile row are still in L2 one tile row later.  Wider chunks are walked down bands of ~4096
// columns, which keeps both properties (a 15360-wide chunk walked row by row re-fetched its halo rows from DRAM: ncu
// showed 1.16x - 2.1x the compulsory reads, against 1.00x - 1.05x at 3840).  Within that, the tiles whose input boxes
// lie entirely inside the cells 1..nx x 1..ny (no halo cell, hence
// no dependence on a preceding halo exchange / reflective boundary) come first, followed by the rim
// tiles; n_interior = how many there are.  Tile (tx,ty) reads the columns 1+tx*tw-lo_x .. 1+(tx+1)*tw-1+hi_x and the
// rows 1+ty*th-lo_y .. 1+(ty+1)*th-1+hi_y.  See "programmatic dependent launch" in common.cuh.
struct TileOrder {
  const int2* table;
  int ntiles;
  int n_interior;
};
TileOrder tile_order_split(int ntx, int nty, int tw, int th, int lo_x, int hi_x, int lo_y, int hi_y, int nx, int ny);
// dep_start for a compute launch: the number of interior tiles when the previous launch was a halo kernel, else 0
// (wait before the first tile).
int dep_start_for(const TileOrder& o);

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// generic-proxy accesses to shared memory (the reads of the tile that used this stage before) are ordered before
// the async-proxy writes of the TMA load issued next
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 2-D tile load: box of `map` with its (dim0, dim1) corner at (x, y) -> dst; completion (bytes) on `bar`
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"((unsigned long long)map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)map) : "memory");
}

// Shared memory of one SM (228 KB on sm_100) against what CPS resident CTAs take: the dynamic bytes, the few static
// words of the tile queue, and the 1 KB the system reserves per CTA.  A kernel whose __launch_bounds__ promises CPS
// CTAs per SM asserts this, so that a tile-shape change cannot silently halve its occupancy.
constexpr bool fits_sm(int dyn_smem_bytes, int ctas_per_sm) {
  return (long long)ctas_per_sm * (dyn_smem_bytes + 128 + 1024) <= 228 * 1024 && dyn_smem_bytes <= 227 * 1024;
}

// Ring of STAGES shared-memory stages, each holding NARR boxes of BW x BH doubles (each box padded to 128 B).
template <int NARR, int BW, int BH, int STAGES>
struct TileRing {
  static constexpr int BOX_BYTES = BW * BH * 8;
  static constexpr int ARR_BYTES = (BOX_BYTES + 127) / 128 * 128;
  static constexpr int STAGE_BYTES = NARR * ARR_BYTES;
  static constexpr int BYTES = STAGES * STAGE_BYTES + STAGES * 8;  // + the full barriers
  static_assert((BW * 8) % 16 == 0, "TMA: inner box extent must be a multiple of 16 bytes");
  static_assert(BW <= 256 && BH <= 256, "TMA: box extents are at most 256");
  unsigned char* base;  // 128-byte aligned
  uint64_t* full;

  __device__ __forceinline__ void init(unsigned char* smem_aligned) {
    base = smem_aligned;
    full = reinterpret_cast<uint64_t*>(smem_aligned + STAGES * STAGE_BYTES);
    if (threadIdx.x == 0 && threadIdx.y == 0) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
      mbar_init_fence();
    }
    __syncthreads();
  }
  __device__ __forceinline__ const double* tile(int stage, int a) const {
    return reinterpret_cast<const double*>(base + stage * STAGE_BYTES + a * ARR_BYTES);
  }
  // a box whose contents are dead, reused as scratch until the stage is issued again (issue() fences the proxies)
  __device__ __forceinline__ double* scratch(int stage, int a) const {
    return reinterpret_cast<double*>(base + stage * STAGE_BYTES + a * ARR_BYTES);
  }
  // one thread: load the boxes with corner (x, y) of all NARR maps into `stage`
  // (x must be even, see the header comment)
  __device__ __forceinline__ void issue(const CUtensorMap* maps, int stage, int x, int y) {
    fence_proxy_async_smem();
    mbar_arrive_expect_tx(&full[stage], (uint32_t)(NARR * BOX_BYTES));
#pragma unroll
    for (int a = 0; a < NARR; ++a)
      tma_load_2d(base + stage * STAGE_BYTES + a * ARR_BYTES, &maps[a], x, y, &full[stage]);
  }
  __device__ __forceinline__ void wait(int stage, uint32_t parity) { mbar_wait(&full[stage], parity); }
};

// Dynamic tile scheduling for the persistent kernels: CTAs draw tickets from a device counter instead of striding
// through the order table.  Why: (1) CTAs that become resident late -- their SM slot was still held by the
// halo-exchange kernel they overlap with (PDL) or by the tail of the previous kernel -- simply take fewer tiles instead
// of finishing late with a full static share; (2) data-dependent tile costs (the zero-numerator / inactive-limiter
// short cuts) balance out.  Tickets follow the order table, so interior tiles still come first and the tiles in
// flight at any moment are still neighbours (L2 reuse of shared halo rows).
// The CTA's leader keeps the latencies off the critical path: at iteration i it issues the TMA loads of a tile whose
// table entry it asked for at i-1 and whose chunk of tickets it drew at least one chunk earlier (an atomicAdd or a
// table load consumed in the iteration that issues it cost ~1 us per tile: measured +50 % on pdv_predict).
// Every launch has its own counter pair {tickets, exits} from a ring (runtime.cu: next_tickets); the last CTA to
// leave zeroes the pair for its next use, so no host-side arithmetic depends on how many tickets were drawn.
struct Tickets {
  unsigned int* ctr;  // [0] next ticket, [1] CTAs that have left; nullptr: static schedule (A/B switch)
};
// Tickets are drawn QCHUNK at a time: one atomicAdd per tile (57 600 tiles of 32x8 cells at 3840^2, 296 CTAs asking every
// ~1 us) saturates the single L2 address and its queueing latency lands on the leader's critical path (measured:
// pdv_predict 0.225 -> 0.288 ms).  A chunk is also a run of horizontally adjacent tiles, which share halo columns.
#ifndef QCHUNK
#define QCHUNK 4
#endif
template <int STAGES>
struct TileQueue {
  unsigned int* ctr;
  int ntiles;
  const int2* order;
  int* s_tile;   // [STAGES] shared: tile index loaded into each ring stage (>= ntiles: none)
  int2* s_xy;    // [STAGES] shared: its coordinates
  // The leader's look-ahead state lives in shared memory too (s_q, 8 words): registers are allocated for every thread
  // of the kernel, and the advection kernels run at 80 registers (3 CTAs/SM) where six more meant spills in the tile
  // loop.  s_q: [0] t_look, [3] cur_lo, [4] cur_hi, [6] turn (static schedule).  What is still in flight when an
  // iteration ends -- the raw return value of the chunk atomic and the table entry of the look-ahead tile -- stays in
  // registers: a store of either would park the leader's warp (in-order issue) for the whole L2 round trip.
  int* s_q;
  unsigned int nxt_raw;  // first ticket of the next chunk: its atomicAdd is in flight / here
  int2 xy_look;          // coordinates of tile t_look: its table load is in flight / here
  __device__ __forceinline__ TileQueue(Tickets t, int n, const int2* o, int* st, int2* sx, int* sq)
      : ctr(t.ctr), ntiles(n), order(o), s_tile(st), s_xy(sx), s_q(sq), nxt_raw((unsigned int)n), xy_look(make_int2(0, 0)) {}
  __device__ __forceinline__ int clampt(unsigned int v) const { return v < (unsigned int)ntiles ? (int)v : ntiles; }
  // The value an atomicAdd returns is NOT touched here (warps issue in order: a compare right behind the atomic would
  // park the leader's warp for the whole L2 round trip); it is stored, and clamped when the chunk becomes current.
  __device__ __forceinline__ unsigned int draw_chunk() {
    if (ctr == nullptr) return (blockIdx.x + (unsigned int)(s_q[6]++) * gridDim.x) * QCHUNK;  // CTA b: chunks b, b+G, ...
    return atomicAdd(ctr, (unsigned int)QCHUNK);
  }
  __device__ __forceinline__ int pop() {
    int lo = s_q[3];
    const int hi = s_q[4];
    if (lo >= hi) {  // the chunk drawn QCHUNK tiles ago becomes current, the next one is asked for now
      lo = clampt(nxt_raw);
      s_q[4] = min(lo + QCHUNK, ntiles);
      if (lo < ntiles) nxt_raw = draw_chunk();  // nothing more to draw after the first miss
      if (lo >= ntiles) { s_q[3] = lo; return ntiles; }
    }
    s_q[3] = lo + 1;
    return lo;
  }
  // The queue is driven by ONE thread that is not the TMA leader (the "scheduler", lane 0 of warp 1): its bookkeeping
  // -- a dependent chain of ~50 single-thread instructions per tile -- then runs next to the leader's fence + nine
  // UTMALDG issues instead of in front of them (in the leader it cost pdv_predict 12 %).  Protocol, per iteration i
  // (ring stage i % STAGES is being computed):
  //   scheduler  step(i % STAGES): table slot i % STAGES := the tile whose loads the leader issues at iteration i+1
  //              (it is computed at iteration i+STAGES); then draws / looks up the tile after that
  //   leader     reads slot (i + STAGES-1) % STAGES -- written by the scheduler during iteration i-1 -- and issues it
  // with a barrier between each write and its reads (every tile loop has one per iteration).
  __device__ __forceinline__ void prime_all() {  // scheduler, once: slots 0..STAGES-1, then the look-ahead
    s_q[6] = 0;
    const int lo = clampt(draw_chunk());
    s_q[3] = lo;
    s_q[4] = min(lo + QCHUNK, ntiles);
    if (lo < ntiles) nxt_raw = draw_chunk();
    int t0[STAGES];
    int2 xy0[STAGES];
#pragma unroll
    for (int s = 0; s < STAGES; ++s) t0[s] = pop();
    const int tl = pop();
#pragma unroll
    for (int s = 0; s < STAGES; ++s) xy0[s] = t0[s] < ntiles ? __ldg(order + t0[s]) : make_int2(0, 0);
    xy_look = tl < ntiles ? __ldg(order + tl) : make_int2(0, 0);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      s_tile[s] = t0[s];
      s_xy[s] = xy0[s];
    }
    s_q[0] = tl;
  }
  __device__ __forceinline__ void step(int slot) {  // scheduler, every iteration
    s_tile[slot] = s_q[0];
    s_xy[slot] = xy_look;  // its table load was issued an iteration ago
    const int tl = pop();
    xy_look = tl < ntiles ? __ldg(order + tl) : make_int2(0, 0);
    s_q[0] = tl;
  }
  // scheduler, after the tile loop: all my draws have returned (the last value is compared here), so the exit count
  // orders after them; the last CTA out re-arms the pair
  __device__ __forceinline__ void leave() {
    if (ctr == nullptr) return;
    if (nxt_raw == 0xffffffffu && s_q[0] < 0) s_tile[0] = (int)nxt_raw;  // never true: waits for the last draw to return
    __threadfence();
    if (atomicAdd(ctr + 1, 1u) == gridDim.x - 1) {
      ctr[0] = 0;
      ctr[1] = 0;
      __threadfence();
    }
  }
};
// Host: the counter pair for the next launch (runtime.cu).
Tickets next_tickets();

// 128-byte alignment of the dynamic shared memory by POINTER arithmetic: a round trip through uintptr_t makes the
// compiler forget that the address is in the shared window, and every tile access then becomes a generic LD.E / ST.E
// with a 64-bit address pair instead of LDS / STS with a 32-bit address and an immediate offset (SASS of the round-1
// kernels: 50 LD.E.64 per tile loop, hundreds of IMAD / LEA of address arithmetic in issue-bound kernels).
__device__ __forceinline__ unsigned char* align128(unsigned char* p) {
  return p + ((128u - (smem_u32(p) & 127u)) & 127u);
}
#endif

}  // namespace clv
