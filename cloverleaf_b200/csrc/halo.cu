// halo.cu -- reflective boundary (update_halo), halo message pack/unpack (the ABI's clover_(un)pack_message_*_c_), the
// halo exchange and scalar reductions that replace clover_exchange / clover_min / clover_sum -- our own kernels over
// peer memory (NVLink / NVSwitch, cudaIpc-mapped exchange blocks), with ncclSend/ncclRecv/ncclAllReduce as the fallback
// transport and NCCL as the bootstrap -- and the two set-up kernels (initialise_chunk, generate_chunk).  fp64 CUDA for
// sm_100a.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <vector>

#include "clover_b200.h"
#include "common.cuh"
#include "lagrange.cuh"

namespace clv {

// from runtime.cu
bool chunk_registered();
int chunk_nx();
int chunk_ny();
const int* chunk_neighbours();
double* chunk_field_host(int f);
void count_copy(long long h2d, long long d2h);
int sm_count();

// Field geometry by id (data.f90:51-66; types as in clover.f90:690-880 / update_halo_kernel_c.c):
// x_inc,y_inc = extra vertices; m = 0 for cell-centred data, 1 otherwise; sx,sy = sign applied when
// reflecting about a left/right resp. bottom/top wall.
struct FieldDesc {
  double* p;
  int x_inc, y_inc, m;
  double sx, sy;
  int offset;  // message offset in doubles (exchange only)
};
struct FieldTable {
  FieldDesc f[15];
  int n;
};
static const struct { int x_inc, y_inc, m; double sx, sy; Kind kind; } kFieldGeom[15] = {
    {0, 0, 0, 1.0, 1.0, CELL},     // density0
    {0, 0, 0, 1.0, 1.0, CELL},     // density1
    {0, 0, 0, 1.0, 1.0, CELL},     // energy0
    {0, 0, 0, 1.0, 1.0, CELL},     // energy1
    {0, 0, 0, 1.0, 1.0, CELL},     // pressure
    {0, 0, 0, 1.0, 1.0, CELL},     // viscosity
    {0, 0, 0, 1.0, 1.0, CELL},     // soundspeed
    {1, 1, 1, -1.0, 1.0, VERTEX},  // xvel0
    {1, 1, 1, -1.0, 1.0, VERTEX},  // xvel1
    {1, 1, 1, 1.0, -1.0, VERTEX},  // yvel0
    {1, 1, 1, 1.0, -1.0, VERTEX},  // yvel1
    {1, 0, 1, -1.0, 1.0, XFACE},   // vol_flux_x
    {0, 1, 1, 1.0, -1.0, YFACE},   // vol_flux_y
    {1, 0, 1, -1.0, 1.0, XFACE},   // mass_flux_x
    {0, 1, 1, 1.0, -1.0, YFACE},   // mass_flux_y
};

// ------------------------------------------------------------------------------------------------
// update_halo_kernel_c.c:86-712 as ONE pass.  The reference does, per field, bottom, top, then left,
// right (the latter over the already reflected halo rows, which is what fills the corners).  Every
// destination cell of that sequence is a pure function of never-written cells:
//     f(jd,kd) = [sx if jd reflected] * [sy if kd reflected] * f(js,ks)
// with js/ks the mirror index when (jd resp. kd) lies beyond an EXTERNAL face, else the index itself
// (halo data received from a neighbour chunk).  Mirror rules:
//     bottom: 1-k <- m+k         top:   ny+y_inc+k <- ny+y_inc+(1-m)-k
//     left:   1-j <- m+j         right: nx+x_inc+j <- nx+x_inc+(1-m)-j
// `part`: 0 = every reflected cell; 1 = only the cells whose mirror source is the chunk's own data (everything except
// the few cells that lie beyond an external face AND in the halo rows / columns of a face that has a neighbour);
// 2 = only those few, whose source is a halo cell delivered by the exchange.  The peer-memory exchange kernel reflects
// part 1 while it waits for its neighbours' strips and part 2 after unpacking them.
// Returns the source address (nullptr: nothing to do) so that callers can batch several loads before the stores.
__device__ __forceinline__ const double* update_halo_source(const FieldDesc& F, int nx, int ny, int pitch, int depth,
                                                            int ext_left, int ext_right, int ext_bottom, int ext_top, int t,
                                                            int part, double*& dst, double& sign);
__device__ __forceinline__ const double* update_halo_cell(const FieldDesc& F, int nx, int ny, int pitch, int ext_left,
                                                          int ext_right, int ext_bottom, int ext_top, int jd, int kd,
                                                          int part, double*& dst, double& sign);
__device__ __forceinline__ void update_halo_item(const FieldDesc& F, int nx, int ny, int pitch, int depth, int ext_left,
                                                 int ext_right, int ext_bottom, int ext_top, int t, int part = 0) {
  double* dst;
  double sign;
  const double* src = update_halo_source(F, nx, ny, pitch, depth, ext_left, ext_right, ext_bottom, ext_top, t, part, dst, sign);
  if (src) *dst = sign * *src;
}
__device__ __forceinline__ const double* update_halo_source(const FieldDesc& F, int nx, int ny, int pitch, int depth,
                                                            int ext_left, int ext_right, int ext_bottom, int ext_top, int t,
                                                            int part, double*& dst, double& sign) {
  const int W = nx + F.x_inc + 2 * depth;  // width of the bottom/top strips (corners included)
  const int H = ny + F.y_inc;              // height of the left/right strips (corners excluded)
  const int n_bt = 2 * depth * W, n_lr = 2 * depth * H;
  if (t >= n_bt + n_lr) return nullptr;
  int jd, kd;
  if (t < n_bt) {
    const int side = t / (depth * W), r = (t % (depth * W)) / W + 1;
    jd = 1 - depth + (t % W);
    kd = side == 0 ? 1 - r : ny + F.y_inc + r;
  } else {
    const int u = t - n_bt;
    const int side = u / (depth * H), r = (u % (depth * H)) / H + 1;
    kd = 1 + (u % H);
    jd = side == 0 ? 1 - r : nx + F.x_inc + r;
  }
  return update_halo_cell(F, nx, ny, pitch, ext_left, ext_right, ext_bottom, ext_top, jd, kd, part, dst, sign);
}
// the same for one halo cell (jd, kd) given directly
__device__ __forceinline__ const double* update_halo_cell(const FieldDesc& F, int nx, int ny, int pitch, int ext_left,
                                                          int ext_right, int ext_bottom, int ext_top, int jd, int kd,
                                                          int part, double*& dst, double& sign) {
  const bool out_x = jd < 1 || jd > nx + F.x_inc, out_y = kd < 1 || kd > ny + F.y_inc;
  const bool refl_x = (jd < 1 && ext_left) || (jd > nx + F.x_inc && ext_right);
  const bool refl_y = (kd < 1 && ext_bottom) || (kd > ny + F.y_inc && ext_top);
  if (!refl_x && !refl_y) return nullptr;
  // the source is a received halo cell iff the cell is outside in a direction that is NOT reflected
  const bool from_halo = (out_x && !refl_x) || (out_y && !refl_y);
  if ((part == 1 && from_halo) || (part == 2 && !from_halo)) return nullptr;
  int js = jd, ks = kd;
  sign = 1.0;
  if (refl_x) {
    js = (jd < 1) ? F.m + (1 - jd) : nx + F.x_inc + (1 - F.m) - (jd - (nx + F.x_inc));
    sign = sign * F.sx;
  }
  if (refl_y) {
    ks = (kd < 1) ? F.m + (1 - kd) : ny + F.y_inc + (1 - F.m) - (kd - (ny + F.y_inc));
    sign = sign * F.sy;
  }
  dst = F.p + idx2(pitch, jd, kd);
  return F.p + idx2(pitch, js, ks);
}
// grid-stride reflection with XU_R independent loads in flight per thread (the strips are short: latency-bound)
// Reflection inside the exchange kernel: one WARP per segment of SEG cells of one halo line (a field's row r beyond
// the bottom / top face, corners included, or its column r beyond the left / right face), SEG/32 independent loads in
// flight per lane.  The segment is decoded once per warp; per cell only an add remains (the flat one-thread-per-ring-
// index form above costs ~100 instructions of integer division per cell, which dominated the exchange kernel).
constexpr int SEG = 256, SEG_U = SEG / 32;
__device__ __forceinline__ void update_halo_range(const FieldTable& T, int nx, int ny, int pitch, int depth, int4 ext,
                                                  int part, int gtid, int gsize) {
  const int lane = gtid & 31, warp = gtid >> 5, nwarps = gsize >> 5;
  const int maxlen = (nx > ny ? nx : ny) + 1 + 2 * depth;
  const int nseg = (maxlen + SEG - 1) / SEG;
  const int lines = T.n * 4 * depth;  // per field: 4 sides x depth lines
  for (int sidx = warp; sidx < lines * nseg; sidx += nwarps) {
    const int line = sidx / nseg, seg = sidx - line * nseg;
    const int f = line / (4 * depth), q = line - f * 4 * depth, side = q / depth, r = q - side * depth + 1;
    const FieldDesc& F = T.f[f];
    // side 0 bottom, 1 top: jd runs over 1-depth .. nx+x_inc+depth; side 2 left, 3 right: kd runs over 1 .. ny+y_inc
    const int len = side < 2 ? nx + F.x_inc + 2 * depth : ny + F.y_inc;
    double v[SEG_U], sg[SEG_U];
    double* dst[SEG_U];
#pragma unroll
    for (int u = 0; u < SEG_U; ++u) {
      const int e = seg * SEG + u * 32 + lane;
      dst[u] = nullptr;
      if (e < len) {
        int jd, kd;
        if (side < 2) { jd = 1 - depth + e; kd = side == 0 ? 1 - r : ny + F.y_inc + r; }
        else          { kd = 1 + e;         jd = side == 2 ? 1 - r : nx + F.x_inc + r; }
        const double* src = update_halo_cell(F, nx, ny, pitch, ext.x, ext.y, ext.z, ext.w, jd, kd, part, dst[u], sg[u]);
        if (src) v[u] = *src;
        else dst[u] = nullptr;
      }
    }
#pragma unroll
    for (int u = 0; u < SEG_U; ++u)
      if (dst[u]) *dst[u] = sg[u] * v[u];
  }
}
__global__ void __launch_bounds__(256)
    update_halo_kernel(FieldTable T, int nx, int ny, int pitch, int depth, int ext_left, int ext_right,
                       int ext_bottom, int ext_top, unsigned long long* trace, int trigger_first) {
  trace_min(trace, 0);
  // trigger_first: the predecessor is reset_field's ring swap, which touches halo rings only and has itself waited for
  // the last compute kernel -- the next compute kernel's interior tiles may start next to it
  if (trigger_first) pdl_trigger();
  pdl_wait();     // the kernel that produced the interior cells has completed
  if (!trigger_first) pdl_trigger();  // the next compute kernel may start its interior tiles (common.cuh: PDL)
  trace_min(trace, 2);
  update_halo_item(T.f[blockIdx.y], nx, ny, pitch, depth, ext_left, ext_right, ext_bottom, ext_top,
                   (int)(blockIdx.x * blockDim.x + threadIdx.x));
  trace_max(trace, 1);
}

// Chunks narrower than depth+1 cells: the mirror source of a depth-2 halo cell can itself be a halo cell that an
// earlier loop of the same call wrote (or has not written yet), so the one-pass composition above does not hold.
// Here the reference's four loops run as four launches in the reference's order (bottom, top, left, right).
__global__ void __launch_bounds__(256)
    update_halo_seq_kernel(FieldTable T, int nx, int ny, int pitch, int depth, int phase) {
  const FieldDesc F = T.f[blockIdx.y];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (phase < 2) {
    const int W = nx + F.x_inc + 2 * depth;
    if (t >= depth * W) return;
    const int k = t / W + 1, j = 1 - depth + (t % W);
    const int kd = phase == 0 ? 1 - k : ny + F.y_inc + k;
    const int ks = phase == 0 ? F.m + k : ny + F.y_inc + (1 - F.m) - k;
    F.p[idx2(pitch, j, kd)] = F.sy * F.p[idx2(pitch, j, ks)];
  } else {
    const int H = ny + F.y_inc + 2 * depth;
    if (t >= depth * H) return;
    const int j = t / H + 1, k = 1 - depth + (t % H);
    const int jd = phase == 2 ? 1 - j : nx + F.x_inc + j;
    const int js = phase == 2 ? F.m + j : nx + F.x_inc + (1 - F.m) - j;
    F.p[idx2(pitch, jd, k)] = F.sx * F.p[idx2(pitch, js, k)];
  }
}

// ------------------------------------------------------------------------------------------------
// pack_kernel_c.c:29-439, all fields of one face in one launch.  face: 0 left, 1 right, 2 bottom, 3 top.
// Message index (0-based): left/right  off + (jj-1) + (k+depth-1)*depth,  k = 1-depth .. ny+y_inc+depth
//                          bottom/top  off + (kk-1) + (j+depth-1)*depth,  j = 1-depth .. nx+x_inc+depth
template <bool UNPACK>
__global__ void __launch_bounds__(256)
    halo_message_kernel(FieldTable T, int nx, int ny, int pitch, int depth, int face, double* __restrict__ buffer) {
  const FieldDesc F = T.f[blockIdx.y];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  int j, k, index;
  if (face < 2) {
    const int span = ny + F.y_inc + 2 * depth;
    if (t >= span * depth) return;
    const int jj = t % depth + 1, kk = t / depth;  // kk = k + depth - 1
    k = kk - depth + 1;
    index = F.offset + (jj - 1) + kk * depth;
    if (face == 0) j = UNPACK ? 1 - jj : 1 + F.x_inc - 1 + jj;
    else           j = UNPACK ? nx + F.x_inc + jj : nx + 1 - jj;
  } else {
    const int span = nx + F.x_inc + 2 * depth;
    if (t >= span * depth) return;
    // consecutive threads walk along j (unit stride in the field); message stride is `depth`
    const int kk = t / span + 1, jx = t % span;  // jx = j + depth - 1
    j = jx - depth + 1;
    index = F.offset + (kk - 1) + jx * depth;
    if (face == 2) k = UNPACK ? 1 - kk : 1 + F.y_inc - 1 + kk;
    else           k = UNPACK ? ny + F.y_inc + kk : ny + 1 - kk;
  }
  if (UNPACK) F.p[idx2(pitch, j, k)] = buffer[index];
  else        buffer[index] = F.p[idx2(pitch, j, k)];
}

static void launch_message(bool unpack, const FieldTable& T, const Grid& g, int depth, int face, double* buffer) {
  const int edge = (face < 2 ? g.ny : g.nx) + 1 + 2 * depth;
  const dim3 grid((unsigned)((edge * depth + 255) / 256), (unsigned)T.n);
  LaunchScope ls(unpack ? "halo_unpack" : "halo_pack");
  if (unpack) halo_message_kernel<true><<<grid, 256, 0, stream()>>>(T, g.nx, g.ny, g.pitch, depth, face, buffer);
  else        halo_message_kernel<false><<<grid, 256, 0, stream()>>>(T, g.nx, g.ny, g.pitch, depth, face, buffer);
}

// ABI pack/unpack of one field: the message buffer is a HOST array handed to MPI by the caller, so
// the packed strip is copied back (pack) or in (unpack) right here, in both residency modes.
static void abi_message(int face, bool unpack, int* xmin, int* xmax, int* ymin, int* ymax, double* field,
                        double* buffer, int* cell, int* vertex, int* xface, int* yface, int* depth_p,
                        int* field_type, int* buffer_offset) {
  const Grid g = grid_of(xmin, xmax, ymin, ymax);
  const int depth = *depth_p, type = *field_type;
  FieldTable T;
  T.n = 1;
  FieldDesc& F = T.f[0];
  Kind kind;
  if (type == *cell) { F.x_inc = 0; F.y_inc = 0; kind = CELL; }
  else if (type == *vertex) { F.x_inc = 1; F.y_inc = 1; kind = VERTEX; }
  else if (type == *xface) { F.x_inc = 1; F.y_inc = 0; kind = XFACE; }
  else if (type == *yface) { F.x_inc = 0; F.y_inc = 1; kind = YFACE; }
  else fatal("pack/unpack: unknown field_type %d", type);
  F.m = 0; F.sx = F.sy = 1.0;
  F.offset = *buffer_offset;
  const size_t lo = (size_t)F.offset;
  const size_t span = (size_t)((face < 2 ? g.ny + F.y_inc : g.nx + F.x_inc) + 2 * depth);
  const size_t hi = lo + span * depth;
  F.p = dev(g, field, kind, unpack ? INOUT_HALO : IN);
  double* dbuf = dev_buffer(buffer, hi, unpack ? IN : 0, lo, hi);
  launch_message(unpack, T, g, depth, face, dbuf);
  if (!unpack) {
    CLV_CUDA(cudaMemcpyAsync(buffer + lo, dbuf + lo, (hi - lo) * sizeof(double), cudaMemcpyDeviceToHost, stream()));
    count_copy(0, (long long)((hi - lo) * sizeof(double)));
    CLV_CUDA(cudaStreamSynchronize(stream()));
  }
  finish();
}

// ------------------------------------------------------------------------------------------------
// NCCL (loaded lazily so that single-GPU runs never need it and so that a process that already
// holds torch's libnccl shares that copy)
struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  double* snd[4] = {};
  double* rcv[4] = {};
  size_t cap[4] = {};
  double* d_scal = nullptr;
} N;

#define CLV_NCCL(call)                                                                          \
  do {                                                                                          \
    ncclResult_t r_ = (call);                                                                   \
    if (r_ != ncclSuccess) fatal("NCCL error at %s:%d: %s", __FILE__, __LINE__, N.GetErrorString(r_)); \
  } while (0)

static void nccl_load() {
  if (N.lib) return;
  N.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!N.lib) N.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!N.lib) N.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!N.lib) fatal("cannot load libnccl.so.2: %s", dlerror());
#define LOAD(field, name)                                        \
  *(void**)(&N.field) = dlsym(N.lib, name);                      \
  if (!N.field) fatal("libnccl lacks %s", name)
  LOAD(GetUniqueId, "ncclGetUniqueId");
  LOAD(CommInitRank, "ncclCommInitRank");
  LOAD(CommDestroy, "ncclCommDestroy");
  LOAD(Send, "ncclSend");
  LOAD(Recv, "ncclRecv");
  LOAD(GroupStart, "ncclGroupStart");
  LOAD(GroupEnd, "ncclGroupEnd");
  LOAD(AllReduce, "ncclAllReduce");
  LOAD(AllGather, "ncclAllGather");
  LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
}

static void ensure_msg_buffers(int face, size_t doubles) {
  if (N.cap[face] >= doubles) return;
  CLV_CUDA(cudaStreamSynchronize(stream()));
  if (N.snd[face]) CLV_CUDA(cudaFree(N.snd[face]));
  if (N.rcv[face]) CLV_CUDA(cudaFree(N.rcv[face]));
  CLV_CUDA(cudaMalloc(&N.snd[face], doubles * sizeof(double)));
  CLV_CUDA(cudaMalloc(&N.rcv[face], doubles * sizeof(double)));
  N.cap[face] = doubles;
}

// ------------------------------------------------------------------------------------------------
// initialise_chunk_kernel_c.c:60-121
__global__ void init_chunk_1d_kernel(int nx, int ny, double min_x, double min_y, double d_x, double d_y,
                                     double* vertexx, double* vertexdx, double* vertexy, double* vertexdy,
                                     double* cellx, double* celldx, double* celly, double* celldy) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;  // t = index - lower bound(-1)
  const int j = t - 1;                                   // Fortran index; x_min = y_min = 1
  if (t <= nx + 4) {
    vertexx[t] = min_x + d_x * (double)(j - 1);
    vertexdx[t] = d_x;
  }
  if (t <= nx + 3) {
    cellx[t] = 0.5 * ((min_x + d_x * (double)(j - 1)) + (min_x + d_x * (double)(j + 1 - 1)));
    celldx[t] = d_x;
  }
  if (t <= ny + 4) {
    vertexy[t] = min_y + d_y * (double)(j - 1);
    vertexdy[t] = d_y;
  }
  if (t <= ny + 3) {
    celly[t] = 0.5 * ((min_y + d_y * (double)(j - 1)) + (min_y + d_y * (double)(j + 1 - 1)));
    celldy[t] = d_y;
  }
}
__global__ void init_chunk_2d_kernel(int nx, int ny, int pitch, double d_x, double d_y, double* volume,
                                     double* xarea, double* yarea) {
  const int j = -XOFF + (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const int k = -1 + (int)(blockIdx.y * blockDim.y + threadIdx.y);
  if (j < -1 || j > nx + 2 || k > ny + 2) return;
  const size_t c = idx2(pitch, j, k);
  volume[c] = d_x * d_y;
  xarea[c] = d_y;  // celldy(k)
  yarea[c] = d_x;  // celldx(j)
}

// generate_chunk_kernel_c.c:70-158, gathered per destination instead of scattered per cell: a cell
// takes the LAST state (2..n) whose geometry test it passes; a vertex takes the velocity of the last
// state that any of its (up to four) adjacent cells passes.
constexpr int MAX_STATES = 24;
struct States {
  int n;
  int geometry[MAX_STATES];
  double density[MAX_STATES], energy[MAX_STATES], xvel[MAX_STATES], yvel[MAX_STATES];
  double xmin[MAX_STATES], xmax[MAX_STATES], ymin[MAX_STATES], ymax[MAX_STATES], radius[MAX_STATES];
  int g_rect, g_circ, g_point;
};
__device__ __forceinline__ bool state_hits(const States& S, int s, int j, int k, int ny_len, const double* vertexx,
                                           const double* vertexy, const double* cellx, const double* celly) {
  if (S.geometry[s] == S.g_rect)
    return vertexx[j + 2] >= S.xmin[s] && vertexx[j + 1] < S.xmax[s] && vertexy[k + 2] >= S.ymin[s] &&
           vertexy[k + 1] < S.ymax[s];
  if (S.geometry[s] == S.g_circ) {
    const double x_cent = S.xmin[s], y_cent = S.ymin[s];
    const double radius = sqrt((cellx[j + 1] - x_cent) * (cellx[j + 1] - x_cent) +
                               (celly[k + 1] - y_cent) * (celly[k + 1] - y_cent));
    return radius <= S.radius[s];
  }
  if (S.geometry[s] == S.g_point)  // :143 reads vertexy at index j
    return vertexx[j + 1] == S.xmin[s] && vertexy[(j + 1 < ny_len) ? j + 1 : ny_len - 1] == S.ymin[s];
  return false;
}
__global__ void generate_chunk_kernel(States S, int nx, int ny, int pitch, const double* __restrict__ vertexx,
                                      const double* __restrict__ vertexy, const double* __restrict__ cellx,
                                      const double* __restrict__ celly, double* __restrict__ density0,
                                      double* __restrict__ energy0, double* __restrict__ xvel0,
                                      double* __restrict__ yvel0) {
  const int j = -XOFF + (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const int k = -1 + (int)(blockIdx.y * blockDim.y + threadIdx.y);
  if (j < -1 || j > nx + 3 || k > ny + 3) return;
  const size_t c = idx2(pitch, j, k);
  if (j <= nx + 2 && k <= ny + 2) {
    double rho = S.density[0], e = S.energy[0];
    for (int s = 1; s < S.n; ++s)
      if (state_hits(S, s, j, k, ny + 5, vertexx, vertexy, cellx, celly)) {
        rho = S.density[s];
        e = (S.geometry[s] == S.g_rect) ? S.energy[s] : S.density[s];  // :134,:145 quirk
      }
    density0[c] = rho;
    energy0[c] = e;
  }
  bool set = (j <= nx + 2 && k <= ny + 2);
  double u = S.xvel[0], v = S.yvel[0];
  for (int s = 1; s < S.n; ++s) {
    bool hit = false;
    for (int ck = k - 1; ck <= k; ++ck)
      for (int cj = j - 1; cj <= j; ++cj)
        if (cj >= -1 && cj <= nx + 2 && ck >= -1 && ck <= ny + 2)
          hit = hit || state_hits(S, s, cj, ck, ny + 5, vertexx, vertexy, cellx, celly);
    if (hit) { u = S.xvel[s]; v = S.yvel[s]; set = true; }
  }
  if (set) { xvel0[c] = u; yvel0[c] = v; }
}

// ------------------------------------------------------------------------------------------------
// Halo exchange over peer memory (NVLink / NVSwitch), no library call on the data path.
//
// Every rank owns one device block  [header | 4 face slots x2 | 4 corner slots x2]  that every rank can address
// (cudaIpc* handles, exchanged once with ncclAllGather).  An exchange is ONE launch and ONE network round trip:
//   put:     the left/right and bottom/top strips of all requested fields (message layout of pack_kernel_c.c,
//            per-field offsets of clover.f90:368-375, own rows/columns only) go STRAIGHT INTO THE NEIGHBOUR's receive
//            slot with ordinary stores through NVLink, and the depth x depth corner blocks go to the four DIAGONAL
//            neighbours.  (The reference gets the corners by running left/right before bottom/top, clover.f90:377-500:
//            the bottom/top strips then carry the halo columns just received.  The value that ends up in a corner cell
//            is the diagonal neighbour's interior cell either way; corner cells next to an external face are
//            rewritten by the reflective boundary that follows, as in the reference.)
//   publish: the last CTA to finish writes the exchange's sequence number into every peer's header (st.release.sys);
//   wait:    until my header shows that number for every face/corner that has a peer (ld.acquire.sys);
//   get:     unpack;  then a grid barrier and the reflective boundary of the external faces (update_halo_kernel).
// Two slots per face/corner, used alternately: a peer can only start writing exchange n+2 after it has received my
// exchange n+1, which I sent after unpacking n (stream order), so slot n&1 is free again by then.
constexpr int P2P_MAX_RANKS = RT_MAX_RANKS;
constexpr int AR_OFF = RT_OFF, AR_SLOT = RT_SLOT;  // all-reduce mailboxes: [parity][sender rank] x {8 values, sequence number}
constexpr int CORNER_SLOT = 512;  // 15 fields x 2x2 doubles
constexpr int P2P_HEADER = AR_OFF + 2 * P2P_MAX_RANKS * AR_SLOT;  // flags at face*64, tickets at 256, mailboxes at 1024
struct PeerInfo {  // what a rank publishes about its block
  cudaIpcMemHandle_t handle;
  unsigned long long off[4];   // byte offset of slot 0 of my face f
  unsigned long long slot[4];  // slot size in bytes
  unsigned long long ok;
  unsigned long long corner_off;  // byte offset of corner slot (corner 0, parity 0); CORNER_SLOT bytes each, [parity][corner]
  int nb[4];                      // chunk_neighbours (chunk ids, -1 = external)
  unsigned long long pad[4];
};
static_assert(sizeof(PeerInfo) == 192, "PeerInfo is exchanged as raw bytes");
struct P2P {
  bool tried = false, on = false;
  unsigned char* mine = nullptr;
  unsigned long long off[4] = {}, slot[4] = {};
  unsigned char* peer[4] = {};            // the block of the neighbour across my face f
  unsigned long long peer_off[4] = {}, peer_slot[4] = {};  // layout of the neighbour's face opposite to f
  unsigned long long corner_off = 0;
  int diag[4] = {-1, -1, -1, -1};         // rank of the diagonal neighbour: 0 bottom-left, 1 bottom-right, 2 top-left, 3 top-right
  unsigned long long diag_corner_off[4] = {};
  unsigned int gen = 0;              // exchange launches so far = sequence number of the current one
  unsigned int arrive_total = 0, barrier_total = 0;  // running targets of the two monotonic CTA tickets in my header
  unsigned char* all[P2P_MAX_RANKS] = {};  // every rank's block (mine included)
  unsigned char** d_all = nullptr;         // the same table on the device
  unsigned long long ar_seq = 0;
  long long bytes_sent = 0;                       // through peer memory
  long long nccl_bytes = 0, nccl_exchanges = 0;  // through the ncclSend/ncclRecv transport
} PP;

// what the last reduction kernel left in pinned memory (see answer_from_fused below)
struct FusedAllreduce {
  bool valid = false;
  int base = 0, n = 0;
  bool is_min = false;
} FA;

struct XArgs {
  int nface;                         // faces that have a neighbour
  int face[4];
  double* fbuf[4];                   // the neighbour's slot for this exchange
  double* fmine[4];                  // my slot
  int ncorner;                       // corners that have a diagonal neighbour
  int corner[4];                     // 0 bottom-left, 1 bottom-right, 2 top-left, 3 top-right (as seen from me)
  double* cbuf[4];
  double* cmine[4];
  int nflag;                         // peers to notify / to wait for (faces first, then corners)
  unsigned long long* flag_out[8];
  unsigned long long* flag_in[8];
  unsigned long long seq;
};

// corner element t (= (kk-1)*depth + jj-1) of field F: the cell I send towards corner c / the halo cell I fill at c
__device__ __forceinline__ void corner_index(const FieldDesc& F, int nx, int ny, int depth, int c, bool unpack, int t,
                                             int& j, int& k) {
  const int jj = t % depth + 1, kk = t / depth + 1;
  const bool left = (c & 1) == 0, bottom = (c & 2) == 0;
  if (unpack) {
    j = left ? 1 - jj : nx + F.x_inc + jj;
    k = bottom ? 1 - kk : ny + F.y_inc + kk;
  } else {
    j = left ? 1 + F.x_inc - 1 + jj : nx + 1 - jj;
    k = bottom ? 1 + F.y_inc - 1 + kk : ny + 1 - kk;
  }
}

// A few CTAs move all strips: one WARP per segment of SEG cells of one strip line (field f, strip row / column r of a
// face), SEG/32 independent loads in flight per lane; the segment is decoded once per warp (the copy is latency-bound:
// a few hundred KB, strided columns of the fields or the peers' freshly written slots).  Message layout = the
// reference's (pack_kernel_c.c): left/right index = off + (r-1) + (k+depth-1)*depth, bottom/top off + (r-1) +
// (j+depth-1)*depth, off = field ordinal * depth * (edge+5); only own rows / columns travel (corners separately).
template <bool UNPACK>
__device__ __forceinline__ void exchange_copy(const FieldTable& T, int nx, int ny, int pitch, int depth, const XArgs& A,
                                              int gtid, int gsize) {
  const int lane = gtid & 31, warp = gtid >> 5, nwarps = gsize >> 5;
  const int maxlen = (nx > ny ? nx : ny) + 1;
  const int nseg = (maxlen + SEG - 1) / SEG;
  const int lines = T.n * depth;  // per face
  for (int sidx = warp; sidx < A.nface * lines * nseg; sidx += nwarps) {
    const int z = sidx / (lines * nseg), w = sidx - z * lines * nseg, line = w / nseg, seg = w - line * nseg;
    const int f = line / depth, r = line - f * depth + 1;
    const int face = A.face[z];
    const FieldDesc& F = T.f[f];
    const int len = face < 2 ? ny + F.y_inc : nx + F.x_inc;  // own rows (left/right) resp. columns (bottom/top)
    const int off = F.offset * depth * ((face < 2 ? ny : nx) + 5) + (r - 1);
    int jfix = 0, kfix = 0;  // the fixed coordinate of this line
    if (face == 0) jfix = UNPACK ? 1 - r : F.x_inc + r;
    if (face == 1) jfix = UNPACK ? nx + F.x_inc + r : nx + 1 - r;
    if (face == 2) kfix = UNPACK ? 1 - r : F.y_inc + r;
    if (face == 3) kfix = UNPACK ? ny + F.y_inc + r : ny + 1 - r;
    double* const slot = UNPACK ? A.fmine[z] : A.fbuf[z];
    double v[SEG_U];
    double* dst[SEG_U];
#pragma unroll
    for (int u = 0; u < SEG_U; ++u) {
      const int e = seg * SEG + u * 32 + lane;  // 0-based position along the face: cell 1+e
      dst[u] = nullptr;
      if (e < len) {
        const int j = face < 2 ? jfix : 1 + e, k = face < 2 ? 1 + e : kfix;
        double* cell = F.p + idx2(pitch, j, k);
        double* msg = slot + off + (e + depth) * depth;  // ((1+e) + depth - 1) * depth
        if (UNPACK) { v[u] = __ldcg(msg); dst[u] = cell; }
        else        { v[u] = *cell;       dst[u] = msg; }
      }
    }
#pragma unroll
    for (int u = 0; u < SEG_U; ++u)
      if (dst[u]) *dst[u] = v[u];
  }
  const int per_corner = depth * depth;
  for (int i = gtid; i < A.ncorner * T.n * per_corner; i += gsize) {
    const int z = i / (T.n * per_corner), r = i - z * (T.n * per_corner), f = r / per_corner, t = r - f * per_corner;
    const FieldDesc& F = T.f[f];
    int j, k;
    corner_index(F, nx, ny, depth, A.corner[z], UNPACK, t, j, k);
    if (UNPACK) F.p[idx2(pitch, j, k)] = __ldcg(A.cmine[z] + f * per_corner + t);
    else        A.cbuf[z][f * per_corner + t] = F.p[idx2(pitch, j, k)];
  }
}
// Grid-wide barrier.  All CTAs of the exchange kernel are resident: it is launched into an otherwise idle stream
// position (the kernels before it never wait for it, and a dependent kernel can only start once every CTA of this
// one has started -- common.cuh: PDL), with at most one CTA per SM.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target, unsigned long long timeout_ns,
                                             double* err, int rank) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned long long t0 = 0;
    unsigned int polls = 0;
    // signed distance: the tickets wrap after 2^32 CTA arrivals
    while ((int)(*(volatile unsigned int*)counter - target) < 0) {
      if ((++polls & 1023u) == 0) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > timeout_ns) {
          err[1] = (double)rank; err[2] = -1.0; err[3] = (double)target; err[4] = (double)*(volatile unsigned int*)counter;
          __threadfence_system();
          err[0] = 2.0;
          __threadfence_system();
          __trap();
        }
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// `counters`: monotonic tickets in my header; arrive_target / barrier_target = their values once every CTA of this
// launch has arrived.  A.seq = number of exchanges so far = the sequence number every rank uses for this exchange.
__global__ void __launch_bounds__(256)
    halo_exchange_kernel(FieldTable T, int nx, int ny, int pitch, int depth, XArgs A, unsigned int* counters,
                         unsigned int arrive_target, unsigned int barrier_target, int4 ext, unsigned long long timeout_ns,
                         double* err, int rank, unsigned long long* trace, int trigger_first) {
  trace_min(trace, 0);
  if (trigger_first) pdl_trigger();  // (behind reset_field's ring swap: see update_halo_kernel)
  pdl_wait();     // the kernel that produced the strips has completed
  if (!trigger_first) pdl_trigger();  // the next compute kernel may start: its interior tiles need none of what follows
  trace_min(trace, 2);  // [2] = work begins (dependency satisfied); [3] = all neighbours' strips have arrived
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  if (A.nflag > 0) {
    exchange_copy<false>(T, nx, ny, pitch, depth, A, gtid, gsize);
    // publish: the last CTA to arrive tells every peer
    __syncthreads();
    trace_max(trace, 4);  // [4] = strips written into the neighbours' slots
    if (threadIdx.x == 0) {
      __threadfence_system();
      if (atomicAdd(counters + 0, 1u) + 1 == arrive_target) {
        // one system-scope fence (it is cumulative over the other CTAs' strips, ordered before it by their fence +
        // ticket), then the flags go out back to back as relaxed stores -- a st.release per flag would serialise one
        // NVLink round trip per neighbour
        __threadfence_system();
        for (int z = 0; z < A.nflag; ++z)
          asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(A.flag_out[z]), "l"(A.seq) : "memory");
      }
    }
  }
  // The reflective boundary of the external faces (update_halo_kernel_c.c) in the same launch.  Almost all of it
  // mirrors the chunk's own cells and is done now, while the neighbours' strips are on their way ...
  const bool reflect = (ext.x | ext.y | ext.z | ext.w) != 0;
  if (reflect) update_halo_range(T, nx, ny, pitch, depth, ext, A.nflag > 0 ? 1 : 0, gtid, gsize);
  if (A.nflag > 0) {
    if (threadIdx.x == 0) {
      for (int z = 0; z < A.nflag; ++z) spin_until_ge(A.flag_in[z], A.seq, timeout_ns, err, 1, rank, z);
      trace_max(trace, 3);
    }
    __syncthreads();
    exchange_copy<true>(T, nx, ny, pitch, depth, A, gtid, gsize);
    trace_max(trace, 5);  // [5] = unpacked
    // ... except the few cells beyond an external face that mirror halo cells the exchange has just delivered (the
    // corner blocks between an external face and a face with a neighbour): after the unpack, hence the barrier
    if (reflect) {
      grid_barrier(counters + 3, barrier_target, timeout_ns, err, rank);
      trace_max(trace, 6);  // [6] = past the grid barrier
      // part 2 lives in the four depth x depth corner blocks only: enumerate those (a scan of the whole ring for them
      // cost 15-20 us of index arithmetic)
      const int per_corner = depth * depth;
      for (int i = gtid; i < T.n * 4 * per_corner; i += gsize) {
        const int f = i / (4 * per_corner), r = i - f * 4 * per_corner, c = r / per_corner, t = r - c * per_corner;
        const FieldDesc& F = T.f[f];
        const int jj = t % depth + 1, kk = t / depth + 1;
        const int jd = (c & 1) ? nx + F.x_inc + jj : 1 - jj, kd = (c & 2) ? ny + F.y_inc + kk : 1 - kk;
        double* dst;
        double sign;
        const double* src = update_halo_cell(F, nx, ny, pitch, ext.x, ext.y, ext.z, ext.w, jd, kd, 2, dst, sign);
        if (src) *dst = sign * *src;
      }
    }
  }
  trace_max(trace, 1);
}

static bool p2p_setup(const Grid& g) {
  if (PP.tried) return PP.on;
  PP.tried = true;
  if (const char* e = getenv("CLOVER_B200_P2P"))
    if (atoi(e) == 0) return false;
  const int* nb = chunk_neighbours();
  unsigned long long bytes = P2P_HEADER;
  for (int f = 0; f < 4; ++f) {
    PP.slot[f] = (unsigned long long)15 * 2 * ((f < 2 ? g.ny : g.nx) + 5) * sizeof(double);
    PP.off[f] = bytes;
    bytes += 2 * PP.slot[f];
  }
  PP.corner_off = bytes;
  bytes += 2 * 4 * CORNER_SLOT;
  CLV_CUDA(cudaMalloc(&PP.mine, bytes));
  CLV_CUDA(cudaMemset(PP.mine, 0, bytes));
  PeerInfo me;
  memset(&me, 0, sizeof(me));
  me.ok = (cudaIpcGetMemHandle(&me.handle, PP.mine) == cudaSuccess) ? 1 : 0;
  (void)cudaGetLastError();
  for (int f = 0; f < 4; ++f) { me.off[f] = PP.off[f]; me.slot[f] = PP.slot[f]; me.nb[f] = nb[f]; }
  me.corner_off = PP.corner_off;
  // publish / collect
  unsigned char* d_all = nullptr;
  CLV_CUDA(cudaMalloc(&d_all, (size_t)(N.nranks + 1) * sizeof(PeerInfo)));
  CLV_CUDA(cudaMemcpyAsync(d_all + (size_t)N.nranks * sizeof(PeerInfo), &me, sizeof(me), cudaMemcpyHostToDevice, stream()));
  CLV_NCCL(N.AllGather(d_all + (size_t)N.nranks * sizeof(PeerInfo), d_all, sizeof(PeerInfo), ncclChar, N.comm, stream()));
  std::vector<PeerInfo> all(N.nranks);
  CLV_CUDA(cudaMemcpyAsync(all.data(), d_all, (size_t)N.nranks * sizeof(PeerInfo), cudaMemcpyDeviceToHost, stream()));
  CLV_CUDA(cudaStreamSynchronize(stream()));
  double ok = (me.ok && N.nranks <= P2P_MAX_RANKS) ? 1.0 : 0.0;
  for (int r = 0; r < N.nranks && ok > 0; ++r) {
    if (r == N.rank) { PP.all[r] = PP.mine; continue; }
    void* ptr = nullptr;
    if (!all[r].ok || cudaIpcOpenMemHandle(&ptr, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      (void)cudaGetLastError();
      ok = 0.0;
      break;
    }
    PP.all[r] = (unsigned char*)ptr;
  }
  for (int f = 0; f < 4 && ok > 0; ++f) {
    if (nb[f] == -1) continue;
    const PeerInfo& q = all[nb[f] - 1];  // rank = chunk - 1 (clover.f90:892)
    PP.peer[f] = PP.all[nb[f] - 1];
    PP.peer_off[f] = q.off[f ^ 1];
    PP.peer_slot[f] = q.slot[f ^ 1];
    if (PP.peer_slot[f] != PP.slot[f]) ok = 0.0;  // both sides of a face see the same edge length
  }
  // diagonal neighbours: the bottom/top neighbour of my left/right neighbour (clover.f90:180-188 numbering)
  for (int c = 0; c < 4 && ok > 0; ++c) {
    const int fx = (c & 1) ? 1 : 0, fy = (c & 2) ? 3 : 2;  // face towards the corner: left/right, bottom/top
    PP.diag[c] = -1;
    if (nb[fx] == -1 || nb[fy] == -1) continue;
    const int via_x = all[nb[fx] - 1].nb[fy], via_y = all[nb[fy] - 1].nb[fx];
    if (via_x == -1 || via_x != via_y) { ok = 0.0; break; }  // not a rectangular decomposition
    PP.diag[c] = via_x - 1;
    PP.diag_corner_off[c] = all[via_x - 1].corner_off;
  }
  if (ok > 0) {
    CLV_CUDA(cudaMalloc(&PP.d_all, sizeof(PP.all)));
    CLV_CUDA(cudaMemcpyAsync(PP.d_all, PP.all, sizeof(PP.all), cudaMemcpyHostToDevice, stream()));
  }
  // every rank must take the same path
  CLV_CUDA(cudaMemcpyAsync(N.d_scal, &ok, sizeof(double), cudaMemcpyHostToDevice, stream()));
  CLV_NCCL(N.AllReduce(N.d_scal, N.d_scal, 1, ncclDouble, ncclMin, N.comm, stream()));
  CLV_CUDA(cudaMemcpyAsync(&ok, N.d_scal, sizeof(double), cudaMemcpyDeviceToHost, stream()));
  CLV_CUDA(cudaStreamSynchronize(stream()));
  CLV_CUDA(cudaFree(d_all));
  PP.on = ok > 0.0;
  if (!PP.on && N.rank == 0)
    fprintf(stderr, "libclover_b200: peer-memory halo exchange unavailable (cudaIpc), using ncclSend/ncclRecv\n");
  return PP.on;
}

void fill_reduce_tail_ranks(ReduceTail& t) {
  if (!PP.on || N.nranks <= 1) return;
  t.all = PP.d_all;
  t.nranks = N.nranks;
  t.rank = N.rank;
  t.ar_seq = ++PP.ar_seq;
}
void note_fused_allreduce(int base, int n, bool is_min, bool across_ranks);

static void p2p_release() {
  for (int r = 0; r < P2P_MAX_RANKS; ++r)
    if (PP.all[r] && PP.all[r] != PP.mine) cudaIpcCloseMemHandle(PP.all[r]);
  if (PP.d_all) cudaFree(PP.d_all);
  if (PP.mine) cudaFree(PP.mine);
  PP = P2P();
}

// All-reduce of up to 8 doubles over peer memory: every rank drops its values, then the sequence number (release,
// system scope), into its mailbox in every rank's block; thread r of the single CTA waits for rank r's mailbox and
// thread 0 folds the values in rank order, so all ranks compute bit-identical results (clover_min: clover.f90:
// 3641-3657 MPI_ALLREDUCE(MIN); clover_sum: :3621-3639 MPI_REDUCE(SUM) to rank 0).  `in`/`out` are pinned, device-visible host memory.
__global__ void __launch_bounds__(P2P_MAX_RANKS)
    p2p_allreduce_kernel(unsigned char** all, int nranks, int rank, const double* in, double* out, int n, int is_min,
                         unsigned long long seq, unsigned long long timeout_ns, double* err) {
  __shared__ double v[P2P_MAX_RANKS][8];
  const int r = threadIdx.x;
  const size_t box = AR_OFF + (size_t)(seq & 1) * P2P_MAX_RANKS * AR_SLOT;
  if (r < nranks) {
    double* dst = reinterpret_cast<double*>(all[r] + box + (size_t)rank * AR_SLOT);
    for (int i = 0; i < n; ++i) dst[i] = in[i];
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst + 8), "l"(seq) : "memory");
    const double* src = reinterpret_cast<const double*>(all[rank] + box + (size_t)r * AR_SLOT);
    spin_until_ge(reinterpret_cast<const unsigned long long*>(src + 8), seq, timeout_ns, err, 3, rank, r);
    for (int i = 0; i < n; ++i) v[r][i] = __ldcg(src + i);
  }
  __syncthreads();
  if (r == 0) {
    for (int i = 0; i < n; ++i) {
      double a = v[0][i];
      for (int q = 1; q < nranks; ++q) a = is_min ? ((v[q][i] < a) ? v[q][i] : a) : a + v[q][i];
      out[i] = a;
    }
    __threadfence_system();
  }
}

// the whole exchange (both phases) + the reflective boundary through peer memory: one launch
static void p2p_exchange(const Grid& g, const HaloArgs& h, const HaloArgs* bc) {
  const int* nb = chunk_neighbours();
  const FieldTable T = [&] {
    // message offsets differ between the phases (edge length); they are recomputed in the kernel from the
    // per-field running sum, so the table carries the offsets of BOTH: offset = index * depth * edge is applied there
    FieldTable t;
    t.n = 0;
    for (int f = 0; f < 15; ++f) {
      if (!h.fields[f]) continue;
      FieldDesc& F = t.f[t.n];
      F.p = dev(g, h.host[f], kFieldGeom[f].kind, INOUT_HALO);
      F.x_inc = kFieldGeom[f].x_inc; F.y_inc = kFieldGeom[f].y_inc; F.m = kFieldGeom[f].m;
      F.sx = kFieldGeom[f].sx; F.sy = kFieldGeom[f].sy;
      F.offset = t.n;  // field ordinal; the kernel scales it by depth*(edge+5) per phase (clover.f90:368-375)
      t.n++;
    }
    return t;
  }();
  if (T.n == 0) return;
  for (int f = 0; f < 4; ++f)
    if ((unsigned long long)15 * 2 * ((f < 2 ? g.ny : g.nx) + 5) * sizeof(double) != PP.slot[f])
      fatal("exchange: chunk %d x %d differs from the one the peer-memory slots were sized for", g.nx, g.ny);
  const unsigned int gen = ++PP.gen;
  XArgs A;
  A.nface = A.ncorner = A.nflag = 0;
  A.seq = gen;
  const size_t par = gen & 1;
  for (int face = 0; face < 4; ++face) {
    if (nb[face] == -1) continue;
    const int i = A.nface++;
    A.face[i] = face;
    A.fbuf[i] = (double*)(PP.peer[face] + PP.peer_off[face] + par * PP.peer_slot[face]);
    A.fmine[i] = (double*)(PP.mine + PP.off[face] + par * PP.slot[face]);
    A.flag_out[A.nflag] = (unsigned long long*)(PP.peer[face] + (face ^ 1) * 64);
    A.flag_in[A.nflag] = (unsigned long long*)(PP.mine + face * 64);
    A.nflag++;
  }
  for (int c = 0; c < 4; ++c) {
    if (PP.diag[c] < 0) continue;
    const int i = A.ncorner++;
    A.corner[i] = c;
    unsigned char* peer = PP.all[PP.diag[c]];
    // I fill the peer's opposite corner (3 - c) and the peer fills my corner c
    A.cbuf[i] = (double*)(peer + PP.diag_corner_off[c] + (par * 4 + (3 - c)) * CORNER_SLOT);
    A.cmine[i] = (double*)(PP.mine + PP.corner_off + (par * 4 + c) * CORNER_SLOT);
    A.flag_out[A.nflag] = (unsigned long long*)(peer + 512 + (3 - c) * 64);
    A.flag_in[A.nflag] = (unsigned long long*)(PP.mine + 512 + c * 64);
    A.nflag++;
  }
  for (int f = 0; f < 15; ++f) {
    if (!h.fields[f]) continue;
    for (int i = 0; i < A.nface; ++i)
      PP.bytes_sent += (long long)h.depth * ((A.face[i] < 2 ? g.ny + kFieldGeom[f].y_inc : g.nx + kFieldGeom[f].x_inc)) * 8;
    PP.bytes_sent += (long long)A.ncorner * h.depth * h.depth * 8;
  }
  // grid: ~4 strip elements per thread, at most one CTA per SM (all resident: the kernel has a grid barrier)
  long long elements = 0;
  for (int i = 0; i < A.nface; ++i) elements += (long long)T.n * h.depth * ((A.face[i] < 2 ? g.ny : g.nx) + 1 + 2 * h.depth);
  int4 ext = make_int4(0, 0, 0, 0);
  if (bc) ext = make_int4(bc->ext[0], bc->ext[1], bc->ext[2], bc->ext[3]);
  const bool reflect = (ext.x | ext.y | ext.z | ext.w) != 0;
  if (reflect) {
    const long long ring = (long long)T.n * (2 * h.depth * (g.nx + 1 + 2 * h.depth) + 2 * h.depth * (g.ny + 1));
    if (ring > elements) elements = ring;
  }
  // Few CTAs on purpose: the exchange runs NEXT TO the interior tiles of the compute kernel that follows it (PDL), and
  // every SM it sits on can hold one compute CTA less while it runs.  $CLOVER_B200_XCTAS overrides (A/B runs).
  static int max_ctas = 0;
  if (!max_ctas) {
    max_ctas = 32;
    if (const char* e = getenv("CLOVER_B200_XCTAS")) max_ctas = atoi(e) > 0 ? atoi(e) : max_ctas;
    if (max_ctas > sm_count()) max_ctas = sm_count();
  }
  int ctas = (int)((elements + 256 * SEG_U - 1) / (256 * SEG_U));
  if (ctas < 1) ctas = 1;
  if (ctas > max_ctas) ctas = max_ctas;
  unsigned int* counters = (unsigned int*)(PP.mine + 256);
  PP.arrive_total += (A.nflag > 0) ? (unsigned)ctas : 0u;
  if (reflect && A.nflag > 0) PP.barrier_total += (unsigned)ctas;
  const int trigger_first = ring_swap_just_launched() ? 1 : 0;
  {
    LaunchScope ls("halo_exchange_p2p");
    launch_pdl(halo_exchange_kernel, dim3((unsigned)ctas), dim3(256), 0, stream(), T, g.nx, g.ny, g.pitch, h.depth, A, counters,
               PP.arrive_total, PP.barrier_total, ext, spin_timeout_ns(), device_error_record(), N.rank, ls.trace,
               trigger_first);
  }
  note_halo_launch();
}

// ---- update_halo and the NCCL exchange on their own (host side) ---------------------------------------
static FieldTable field_table(const Grid& g, const HaloArgs& h, int depth, int edge) {
  FieldTable T;
  T.n = 0;
  int off = 0;
  for (int f = 0; f < 15; ++f) {
    if (!h.fields[f]) continue;
    FieldDesc& F = T.f[T.n++];
    F.p = dev(g, h.host[f], kFieldGeom[f].kind, INOUT_HALO);
    F.x_inc = kFieldGeom[f].x_inc; F.y_inc = kFieldGeom[f].y_inc; F.m = kFieldGeom[f].m;
    F.sx = kFieldGeom[f].sx; F.sy = kFieldGeom[f].sy;
    F.offset = off;  // clover.f90:368-375: per-field offsets are running sums of depth*(edge+5)
    off += depth * edge;
  }
  return T;
}

void note_fused_allreduce(int base, int n, bool is_min, bool across_ranks) {
  FA.valid = across_ranks;
  FA.base = base;
  FA.n = n;
  FA.is_min = is_min;
}

void run_update_halo(const Grid& g, const HaloArgs& h) {
  const FieldTable T = field_table(g, h, h.depth, 0);
  if (T.n > 0 && (h.ext[0] || h.ext[1] || h.ext[2] || h.ext[3])) {
    if (g.nx < h.depth + 1 || g.ny < h.depth + 1) {
      for (int phase = 0; phase < 4; ++phase) {
        if (!h.ext[phase < 2 ? 2 + phase : phase - 2]) continue;  // ext = {left, right, bottom, top}
        const int strip = h.depth * ((phase < 2 ? g.nx : g.ny) + 1 + 2 * h.depth);
        const dim3 grid((unsigned)((strip + 255) / 256), (unsigned)T.n);
        LaunchScope ls("update_halo_seq");
        update_halo_seq_kernel<<<grid, 256, 0, stream()>>>(T, g.nx, g.ny, g.pitch, h.depth, phase);
      }
      return;
    }
    const int ring = 2 * h.depth * (g.nx + 1 + 2 * h.depth) + 2 * h.depth * (g.ny + 1);
    const dim3 grid((unsigned)((ring + 255) / 256), (unsigned)T.n);
    const int trigger_first = ring_swap_just_launched() ? 1 : 0;
    {
      LaunchScope ls("update_halo");
      launch_pdl(update_halo_kernel, grid, dim3(256), 0, stream(), T, g.nx, g.ny, g.pitch, h.depth, h.ext[0], h.ext[1],
                 h.ext[2], h.ext[3], ls.trace, trigger_first);
    }
    note_halo_launch();
  }
}

void run_exchange(const Grid& g, const HaloArgs& h) {
  const int* nb = chunk_neighbours();
  const int depth = h.depth;
  if (p2p_setup(g)) {
    p2p_exchange(g, h, nullptr);
    return;
  }
  PP.nccl_exchanges++;
  for (int phase = 0; phase < 2; ++phase) {
    const int fa = phase * 2, fb = fa + 1;
    if (nb[fa] == -1 && nb[fb] == -1) continue;
    const int edge = (phase == 0 ? g.ny : g.nx) + 5;
    const FieldTable T = field_table(g, h, depth, edge);
    if (T.n == 0) return;
    const size_t total = (size_t)T.n * depth * edge;
    for (int face = fa; face <= fb; ++face) {
      if (nb[face] == -1) continue;
      ensure_msg_buffers(face, (size_t)15 * 2 * edge);
      launch_message(false, T, g, depth, face, N.snd[face]);
    }
    CLV_NCCL(N.GroupStart());
    for (int face = fa; face <= fb; ++face) {
      if (nb[face] == -1) continue;
      const int peer = nb[face] - 1;  // rank = chunk - 1 (clover.f90:892)
      CLV_NCCL(N.Send(N.snd[face], total, ncclDouble, peer, N.comm, stream()));
      PP.nccl_bytes += (long long)(total * sizeof(double));
      CLV_NCCL(N.Recv(N.rcv[face], total, ncclDouble, peer, N.comm, stream()));
    }
    CLV_NCCL(N.GroupEnd());
    for (int face = fa; face <= fb; ++face) {
      if (nb[face] == -1) continue;
      launch_message(true, T, g, depth, face, N.rcv[face]);
    }
  }
}

// clover_exchange followed by update_halo_kernel with the same field list (update_halo.f90:39-113): one launch when
// the peer-memory transport is up and the chunk is wide enough for the one-pass reflection.
void run_exchange_then_halo(const Grid& g, const HaloArgs* ex, const HaloArgs* uh) {
  if (ex && uh && g.nx >= uh->depth + 1 && g.ny >= uh->depth + 1 && ex->depth == uh->depth && p2p_setup(g)) {
    bool same = true;
    for (int f = 0; f < 15; ++f) same = same && (ex->fields[f] == uh->fields[f]) && (!ex->fields[f] || ex->host[f] == uh->host[f]);
    const int* nb = chunk_neighbours();
    const bool any_nb = nb[0] != -1 || nb[1] != -1 || nb[2] != -1 || nb[3] != -1;
    if (same && any_nb) {
      p2p_exchange(g, *ex, uh);
      return;
    }
  }
  if (ex) run_exchange(g, *ex);
  if (uh) run_update_halo(g, *uh);
}

}  // namespace clv

using namespace clv;

extern "C" {

void update_halo_kernel_c_(int* xmin, int* xmax, int* ymin, int* ymax, int* chunk_neighbours,
                           int* tile_neighbours, double* density0, double* energy0, double* pressure,
                           double* viscosity, double* soundspeed, double* density1, double* energy1,
                           double* xvel0, double* yvel0, double* xvel1, double* yvel1, double* vol_flux_x,
                           double* vol_flux_y, double* mass_flux_x, double* mass_flux_y, int* fields,
                           int* depth_p) {
  // Op::a[0..14] = the 15 fields in field-id order (data.f90:51-66); fields = mask; iv = {depth, ext[4]}
  Op op;
  op.kind = OP_UPDATE_HALO;
  const Grid g = op.g = grid_of_noflush(xmin, xmax, ymin, ymax);
  const int depth = op.iv[0] = *depth_p;
  if (depth < 1 || depth > 2) fatal("update_halo: depth %d", depth);
  for (int f = 0; f < 4; ++f) op.iv[1 + f] = (chunk_neighbours[f] == -1 && tile_neighbours[f] == -1);
  double* host[15] = {density0, density1, energy0, energy1, pressure, viscosity, soundspeed, xvel0,
                      xvel1, yvel0, yvel1, vol_flux_x, vol_flux_y, mass_flux_x, mass_flux_y};
  HaloArgs h;
  for (int f = 0; f < 15; ++f) {
    op.a[f] = h.host[f] = host[f];
    op.fields[f] = h.fields[f] = (fields[f] == 1);
    if (h.fields[f]) { op.reads({host[f]}); op.writes({host[f]}); }
  }
  op.na = 15;
  h.depth = depth;
  for (int f = 0; f < 4; ++f) h.ext[f] = op.iv[1 + f];
  op.run = [=] { run_update_halo(g, h); };
  submit(std::move(op));
}

#define CLV_PACK_ENTRY(name, face, unpack)                                                          \
  void name(int* xmin, int* xmax, int* ymin, int* ymax, double* field, double* buffer, int* c, int* v, \
            int* xf, int* yf, int* depth, int* field_type, int* buffer_offset) {                     \
    abi_message(face, unpack, xmin, xmax, ymin, ymax, field, buffer, c, v, xf, yf, depth, field_type, \
                buffer_offset);                                                                      \
  }
CLV_PACK_ENTRY(clover_pack_message_left_c_, 0, false)
CLV_PACK_ENTRY(clover_unpack_message_left_c_, 0, true)
CLV_PACK_ENTRY(clover_pack_message_right_c_, 1, false)
CLV_PACK_ENTRY(clover_unpack_message_right_c_, 1, true)
CLV_PACK_ENTRY(clover_pack_message_bottom_c_, 2, false)
CLV_PACK_ENTRY(clover_unpack_message_bottom_c_, 2, true)
CLV_PACK_ENTRY(clover_pack_message_top_c_, 3, false)
CLV_PACK_ENTRY(clover_unpack_message_top_c_, 3, true)

// ---- communicator + exchange ---------------------------------------------------------------------
void clover_b200_comm_get_unique_id_(char* id128) {
  ensure_init();
  nccl_load();
  ncclUniqueId id;
  CLV_NCCL(N.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
}

void clover_b200_comm_init_(int* nranks, int* rank, char* id128) {
  ensure_init();
  nccl_load();
  if (N.comm) fatal("comm_init called twice");
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  N.nranks = *nranks;
  N.rank = *rank;
  CLV_NCCL(N.CommInitRank(&N.comm, N.nranks, id, N.rank));
  CLV_CUDA(cudaMalloc(&N.d_scal, 16 * sizeof(double)));
}

void clover_b200_comm_finalize_internal() {
  p2p_release();
  if (N.comm) {
    N.CommDestroy(N.comm);
    N.comm = nullptr;
  }
  for (int f = 0; f < 4; ++f) {
    if (N.snd[f]) cudaFree(N.snd[f]);
    if (N.rcv[f]) cudaFree(N.rcv[f]);
    N.snd[f] = N.rcv[f] = nullptr;
    N.cap[f] = 0;
  }
  if (N.d_scal) cudaFree(N.d_scal);
  N.d_scal = nullptr;
  N.nranks = 1;
  N.rank = 0;
}

void clover_b200_exchange_(int* fields, int* depth_p) {
  ensure_init();
  if (!chunk_registered()) fatal("exchange before register_chunk");
  const int* nb = chunk_neighbours();
  if (nb[0] == -1 && nb[1] == -1 && nb[2] == -1 && nb[3] == -1) return;
  if (!N.comm) fatal("exchange with neighbours but no communicator (call clover_b200_comm_init_)");
  // Op::a[0..14] = the registered chunk's 15 fields; fields = mask; iv[0] = depth
  Op op;
  op.kind = OP_EXCHANGE;
  int one = 1, nx = chunk_nx(), ny = chunk_ny();
  const Grid g = op.g = grid_of_noflush(&one, &nx, &one, &ny);
  const int depth = op.iv[0] = *depth_p;
  HaloArgs h;
  for (int f = 0; f < 15; ++f) {
    op.a[f] = h.host[f] = chunk_field_host(f);
    op.fields[f] = h.fields[f] = (fields[f] == 1);
    if (h.fields[f]) { op.reads({h.host[f]}); op.writes({h.host[f]}); }
  }
  op.na = 15;
  h.depth = depth;
  op.run = [=] { run_exchange(g, h); };
  submit(std::move(op));
}

// The reduction kernels (calc_dt / the fused timestep launch, field_summary) fold across ranks themselves when the
// peer-memory transport is up (lagrange.cuh: block_reduce_publish).  What they leave in pinned memory -- this rank's
// values at [base+32..], the all-rank result at [base..] -- is remembered here, and the clover_min / clover_sum call
// the driver makes next (timestep.f90:90, field_summary.f90:103-107) is answered from it without a launch.

static bool answer_from_fused(double* values, int n, bool is_min) {
  if (!FA.valid || FA.is_min != is_min || FA.n != n) return false;
  FA.valid = false;
  const double* h = host_scalars();
  if (is_min) {
    // min over ranks of min(local_r, c) with a rank-uniform c (timestep.f90:86-89: dtold*dtrise, dtmax) equals
    // min(global, value); anything else is not the call sequence this answer is valid for
    if (!(values[0] <= h[FA.base + 32]))
      fatal("clover_b200_min_: the value passed (%.17g) exceeds the dt_min_val calc_dt returned (%.17g)", values[0],
            h[FA.base + 32]);
    if (h[FA.base] < values[0]) values[0] = h[FA.base];
    return true;
  }
  for (int i = 0; i < n; ++i)
    if (values[i] != h[FA.base + 32 + i])
      fatal("clover_b200_sum_: the values passed are not the ones field_summary_kernel_c_ returned");
  for (int i = 0; i < n; ++i) values[i] = h[FA.base + i];
  return true;
}

static void allreduce_host(double* values, int n, ncclRedOp_t op) {
  ensure_init();
  if (!N.comm || N.nranks == 1) {
    flush_deferred();
    return;
  }
  if (n > 16) fatal("allreduce of %d values (max 16)", n);
  // Answered from what the reduction kernel left behind when nothing has been recorded since (the call the driver
  // makes right after calc_dt / field_summary): no launch, and no flush either -- a flush here would issue the
  // held-back viscosity halo update on its own instead of merged with the pressure exchange (fuse.cu).
  if (deferred_queue_empty() && answer_from_fused(values, n, op == ncclMin)) return;
  FA.valid = false;
  flush_deferred();
  if (n <= 8 && chunk_registered()) {
    int one = 1, nx = chunk_nx(), ny = chunk_ny();
    if (p2p_setup(grid_of_noflush(&one, &nx, &one, &ny))) {
      double* h = host_scalars();  // pinned + mapped: [16..23] in, [24..31] out
      for (int i = 0; i < n; ++i) h[16 + i] = values[i];
      {
        LaunchScope ls("allreduce_p2p");
        p2p_allreduce_kernel<<<1, P2P_MAX_RANKS, 0, stream()>>>(PP.d_all, N.nranks, N.rank, h + 16, h + 24, n,
                                                               op == ncclMin ? 1 : 0, ++PP.ar_seq, spin_timeout_ns(),
                                                               device_error_record());
      }
      CLV_CUDA(cudaStreamSynchronize(stream()));
      for (int i = 0; i < n; ++i) values[i] = h[24 + i];
      return;
    }
  }
  CLV_CUDA(cudaMemcpyAsync(N.d_scal, values, n * sizeof(double), cudaMemcpyHostToDevice, stream()));
  CLV_NCCL(N.AllReduce(N.d_scal, N.d_scal, (size_t)n, ncclDouble, op, N.comm, stream()));
  CLV_CUDA(cudaMemcpyAsync(values, N.d_scal, n * sizeof(double), cudaMemcpyDeviceToHost, stream()));
  CLV_CUDA(cudaStreamSynchronize(stream()));
}
// *p2p = 1: halos and reductions travel through peer memory; 0: ncclSend/ncclRecv/ncclAllReduce (or one rank only)
void clover_b200_transport_(int* p2p) { *p2p = (PP.on && N.nranks > 1) ? 1 : 0; }
void clover_b200_halo_bytes_(long long* bytes, long long* exchanges) {
  flush_deferred();
  *bytes = PP.bytes_sent + PP.nccl_bytes;
  *exchanges = (long long)PP.gen + PP.nccl_exchanges;
}
void clover_b200_min_(double* value) { allreduce_host(value, 1, ncclMin); }
void clover_b200_sum_(double* values, int* n) { allreduce_host(values, *n, ncclSum); }

// ---- set-up kernels -------------------------------------------------------------------------------
void initialise_chunk_kernel_c_(int* xmin, int* xmax, int* ymin, int* ymax, double* min_x, double* min_y,
                                double* dx, double* dy, double* vertexx, double* vertexdx, double* vertexy,
                                double* vertexdy, double* cellx, double* celldx, double* celly,
                                double* celldy, double* volume, double* xarea, double* yarea) {
  const Grid g = grid_of(xmin, xmax, ymin, ymax);
  double* vx = dev(g, vertexx, X1D_VERT, OUT);
  double* vdx = dev(g, vertexdx, X1D_VERT, OUT);
  double* vy = dev(g, vertexy, Y1D_VERT, OUT);
  double* vdy = dev(g, vertexdy, Y1D_VERT, OUT);
  double* cx = dev(g, cellx, X1D_CELL, OUT);
  double* cdx = dev(g, celldx, X1D_CELL, OUT);
  double* cy = dev(g, celly, Y1D_CELL, OUT);
  double* cdy = dev(g, celldy, Y1D_CELL, OUT);
  double* vol = dev(g, volume, CELL, OUT);
  double* xa = dev(g, xarea, XFACE, OUT);
  double* ya = dev(g, yarea, YFACE, OUT);
  const int n1 = (g.nx > g.ny ? g.nx : g.ny) + 5;
  {
    LaunchScope ls("initialise_chunk_1d");
    init_chunk_1d_kernel<<<(n1 + 255) / 256, 256, 0, stream()>>>(g.nx, g.ny, *min_x, *min_y, *dx, *dy, vx, vdx, vy,
                                                                 vdy, cx, cdx, cy, cdy);
  }
  {
    const dim3 block(32, 8), grid((unsigned)((g.nx + 3 + XOFF + 32) / 32), (unsigned)((g.ny + 4 + 7) / 8));
    LaunchScope ls("initialise_chunk_2d");
    init_chunk_2d_kernel<<<grid, block, 0, stream()>>>(g.nx, g.ny, g.pitch, *dx, *dy, vol, xa, ya);
  }
  // The eight 1-D geometry arrays never change again and host code reads them (visit.f90:127,131 writes
  // vertexx/vertexy into the VTK files): they always come back to the host, resident mode or not.  The three 2-D
  // ones (volume, xarea, yarea) stay on the device; clover_b200_download_ brings one back on request.
  if (is_resident()) {
    for (double* a : {vertexx, vertexdx, vertexy, vertexdy, cellx, celldx, celly, celldy}) clover_b200_download_(a);
  }
  finish();
}

void generate_chunk_kernel_c_(int* xmin, int* xmax, int* ymin, int* ymax, double* vertexx, double* vertexy,
                              double* cellx, double* celly, double* density0, double* energy0,
                              double* xvel0, double* yvel0, int* number_of_states, double* state_density,
                              double* state_energy, double* state_xvel, double* state_yvel,
                              double* state_xmin, double* state_xmax, double* state_ymin,
                              double* state_ymax, double* state_radius, int* state_geometry, int* g_rect,
                              int* g_circ, int* g_point) {
  const Grid g = grid_of(xmin, xmax, ymin, ymax);
  States S;
  S.n = *number_of_states;
  if (S.n < 1 || S.n > MAX_STATES) fatal("generate_chunk: %d states (max %d)", S.n, MAX_STATES);
  for (int s = 0; s < S.n; ++s) {
    S.geometry[s] = state_geometry[s];
    S.density[s] = state_density[s]; S.energy[s] = state_energy[s];
    S.xvel[s] = state_xvel[s]; S.yvel[s] = state_yvel[s];
    S.xmin[s] = state_xmin[s]; S.xmax[s] = state_xmax[s];
    S.ymin[s] = state_ymin[s]; S.ymax[s] = state_ymax[s];
    S.radius[s] = state_radius[s];
  }
  S.g_rect = *g_rect; S.g_circ = *g_circ; S.g_point = *g_point;
  const double* vx = dev(g, vertexx, X1D_VERT, IN);
  const double* vy = dev(g, vertexy, Y1D_VERT, IN);
  const double* cx = dev(g, cellx, X1D_CELL, IN);
  const double* cy = dev(g, celly, Y1D_CELL, IN);
  double* d0 = dev(g, density0, CELL, OUT);
  double* e0 = dev(g, energy0, CELL, OUT);
  double* x0 = dev(g, xvel0, VERTEX, OUT);
  double* y0 = dev(g, yvel0, VERTEX, OUT);
  {
    const dim3 block(32, 8), grid((unsigned)((g.nx + 4 + XOFF + 32) / 32), (unsigned)((g.ny + 5 + 7) / 8));
    LaunchScope ls("generate_chunk");
    generate_chunk_kernel<<<grid, block, 0, stream()>>>(S, g.nx, g.ny, g.pitch, vx, vy, cx, cy, d0, e0, x0, y0);
  }
  finish();
}

}  // extern "C"
