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

// Host: device table of the ntx*nty tile coordinates of a grid of tw-wide tiles, in the order the persistent CTAs walk
// them (CTA b takes entries b, b+G, b+2G, ...; the kernels fetch an entry one iteration before they need it); cached
// per shape.  Chunks up to ~4096 cells wide are walked row by row: the CTAs that run at the same time then stream long
// contiguous row segments and the halo rows shared with the next tile row are still in L2 one tile row later.  Wider
// chunks are walked down bands of ~4096 columns, which keeps both properties (a 15360-wide chunk walked row by row
// re-fetched its halo rows from DRAM: ncu showed 1.16x - 2.1x the compulsory reads, against 1.00x - 1.05x at 3840).
const int2* tile_order(int ntx, int nty, int tw);
// The same table with the tiles whose input boxes lie entirely inside the cells 1..nx x 1..ny (no halo cell, hence
// no dependence on a preceding halo exchange / reflective boundary) in front, banded as above, followed by the rim
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

// Dynamic tile scheduling for the persistent kernels: CTAs draw tickets from a device counter (one atomicAdd per
// tile, by the CTA's leader, STAGES-1 tiles ahead of the tile being computed) instead of striding through the table.
// Why: (1) CTAs that become resident late -- their SM slot was still held by the halo-exchange kernel they overlap
// with (PDL) or by the tail of the previous kernel -- simply take fewer tiles instead of finishing late with a full
// static share; (2) data-dependent tile costs (the zero-numerator / inactive-limiter short cuts) balance out.
// The counter is never reset: the host passes `base` = its value before this launch; every CTA draws exactly one
// ticket past the end, so a launch advances it by ntiles + gridDim.x (runtime.cu: next_tickets).  Tickets follow the
// order table, so interior tiles still come first and tiles in flight at the same time are still neighbours.
struct Tickets {
  unsigned int* counter;
  unsigned int base;
};
template <int STAGES>
struct TileQueue {
  Tickets tk;
  int ntiles;
  const int2* order;
  bool exhausted;
  int* s_tile;   // [STAGES] shared: tile index loaded into each ring stage (>= ntiles: none)
  int2* s_xy;    // [STAGES] shared: its coordinates
  __device__ __forceinline__ TileQueue(Tickets t, int n, const int2* o, int* st, int2* sx)
      : tk(t), ntiles(n), order(o), exhausted(false), s_tile(st), s_xy(sx) {}
  // leader only: draw the next ticket for ring slot `slot`; true (+ index, coordinates) if it is a tile
  __device__ __forceinline__ bool draw(int slot, int& t, int2& xy) {
    t = ntiles;
    if (!exhausted) {
      const unsigned int v = atomicAdd(tk.counter, 1u) - tk.base;
      if (v < (unsigned int)ntiles) t = (int)v;
      else exhausted = true;
    }
    s_tile[slot] = t;
    if (t >= ntiles) return false;
    xy = __ldg(order + t);
    s_xy[slot] = xy;
    return true;
  }
};
// Host: ticket counter + base for a launch of `ctas` persistent CTAs over `ntiles` tiles (runtime.cu).
Tickets next_tickets(int ntiles, int ctas);

__device__ __forceinline__ unsigned char* align128(unsigned char* p) {
  return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 127) & ~(uintptr_t)127);
}
#endif

}  // namespace clv
