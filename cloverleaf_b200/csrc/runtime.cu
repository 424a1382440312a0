// runtime.cu -- device residency, streams, bookkeeping and the non-kernel extension entry points
// of libclover_b200.so.  See include/clover_b200.h for the ABI and common.cuh for the layout.
#include <sys/time.h>

#include <cstdarg>
#include <cstring>
#include <map>
#include <tuple>
#include <string>
#include <unordered_map>
#include <vector>

#include "clover_b200.h"
#include "common.cuh"
#include "tile_order.h"
#include "tma.cuh"

namespace clv {

[[noreturn]] void fatal(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "libclover_b200: fatal: ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
  abort();
}

namespace {

struct Entry {
  double* d = nullptr;
  double* alt = nullptr;
  Kind kind = CELL;
  int nx = 0, ny = 0;
  size_t doubles = 0;
  const double* lazy_src = nullptr;  // pending lazy copy: my update range := that of lazy_src's mirror
  const double* lazy_src2 = nullptr; // non-null: pending lazy soundspeed: my interior := c(lazy_src = density, lazy_src2 = energy)
};

struct Buffer {
  double* d = nullptr;
  size_t doubles = 0;
};

struct Prof {
  double ms = 0;
  long long calls = 0;
};

struct Runtime {
  bool ready = false;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool resident = true;
  std::unordered_map<const void*, Entry> arrays;
  std::unordered_map<const void*, Buffer> buffers;
  std::multimap<size_t, double*> pool;   // parked device buffers by size in doubles (see alloc_zeroed)
  std::vector<const void*> pending_out;  // non-resident mode: arrays to download in finish()
  double* h_scalars = nullptr;           // pinned + mapped
  double* d_partials = nullptr;
  size_t partials_doubles = 0;
  unsigned int* d_ticket = nullptr;
  long long launches = 0;
  long long h2d = 0, d2h = 0;
  bool profiling = false;
  std::vector<Op> queue;  // deferred calls (resident mode)
  bool draining = false;
  bool fuse = true;
  bool tma = true;
  bool overlap = false;  // side-stream overlap of the viscosity exchange: superseded by PDL ($CLOVER_B200_OVERLAP=1 re-enables)
  cudaStream_t side = nullptr;   // second stream: a halo exchange that nothing waits for yet (see side_begin)
  cudaStream_t cur = nullptr;    // non-null while work is being issued to the side stream
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool side_pending = false;
  int sms = 148;
  int lazy_pending = 0;   // number of entries with lazy_src set
  bool pdl = true;        // programmatic dependent launch ($CLOVER_B200_PDL=0 disables)
  bool split = true;      // interior tiles before the halo wait ($CLOVER_B200_SPLIT=0: wait before the first tile)
  bool halo_noted = false;  // the last launch was an exchange / update_halo kernel
  bool ring_swap_noted = false;  // the last launch was reset_field's ring swap
  unsigned int* d_tickets = nullptr;  // ring of {tickets, exits} pairs (next_tickets)
  unsigned int ticket_turn = 0;
  bool trace_on = false;                 // in-situ timeline (clover_b200_trace_)
  unsigned long long* d_trace = nullptr;  // 4 stamps per launch
  std::vector<const char*> trace_names;
  unsigned long long* cur_trace = nullptr;  // slot of the launch whose LaunchScope is open
  unsigned long long scalar_seq = 0;               // sequence number of the reduction results in h_scalars
  unsigned long long spin_timeout_ns = 20000000000ull;  // device-side waits for other GPUs ($CLOVER_B200_SPIN_TIMEOUT_MS)
  std::map<std::string, Prof> prof;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

constexpr size_t TRACE_CAP = 1 << 16;
Runtime R;

void drop_tensor_maps();  // defined with the tensor-map cache below

bool is_2d(Kind k) { return k <= YFACE; }
int host_row(Kind k, int nx) { return (k == CELL || k == YFACE) ? nx + 4 : nx + 5; }
int host_rows(Kind k, int ny) { return (k == CELL || k == XFACE) ? ny + 4 : ny + 5; }
size_t len_1d(Kind k, int nx, int ny) {
  switch (k) {
    case X1D_CELL: return nx + 4;
    case X1D_VERT: return nx + 5;
    case Y1D_CELL: return ny + 4;
    default: return ny + 5;
  }
}

void upload(const Entry& e, const double* host) {
  if (is_2d(e.kind)) {
    const int pitch = pitch_for(e.nx);
    const size_t w = (size_t)host_row(e.kind, e.nx) * sizeof(double);
    const size_t h = host_rows(e.kind, e.ny);
    CLV_CUDA(cudaMemcpy2DAsync(e.d + (XOFF - 1), (size_t)pitch * sizeof(double), host, w, w, h,
                               cudaMemcpyHostToDevice, R.stream));
    R.h2d += (long long)(w * h);
  } else {
    const size_t n = len_1d(e.kind, e.nx, e.ny) * sizeof(double);
    CLV_CUDA(cudaMemcpyAsync(e.d, host, n, cudaMemcpyHostToDevice, R.stream));
    R.h2d += (long long)n;
  }
}

// First sight of an array that the call is about to overwrite on its whole update range (OUT_FULL): only the cells
// outside that range -- two rows below and above, two columns left and right -- carry information from the host.
// Update ranges: cell 1..nx x 1..ny, vertex 1..nx+1 x 1..ny+1, x-face 1..nx+1 x 1..ny, y-face 1..nx x 1..ny+1.
void upload_ring(const Entry& e, const double* host) {
  const int pitch = pitch_for(e.nx);
  const int roww = host_row(e.kind, e.nx), rows = host_rows(e.kind, e.ny);
  const int nj = e.nx + ((e.kind == VERTEX || e.kind == XFACE) ? 1 : 0);
  const int nk = e.ny + ((e.kind == VERTEX || e.kind == YFACE) ? 1 : 0);
  const int jmax = roww - 2, kmax = rows - 2;  // last Fortran index of the extent (lower bound -1)
  const size_t sp = (size_t)roww * sizeof(double), dp = (size_t)pitch * sizeof(double);
  auto h = [&](int j, int k) { return host + (size_t)(k + 1) * roww + (j + 1); };
  auto d = [&](int j, int k) { return e.d + idx2(pitch, j, k); };
  CLV_CUDA(cudaMemcpy2DAsync(d(-1, -1), dp, h(-1, -1), sp, sp, 2, cudaMemcpyHostToDevice, R.stream));
  CLV_CUDA(cudaMemcpy2DAsync(d(-1, nk + 1), dp, h(-1, nk + 1), sp, sp, kmax - nk, cudaMemcpyHostToDevice, R.stream));
  CLV_CUDA(cudaMemcpy2DAsync(d(-1, 1), dp, h(-1, 1), sp, 2 * sizeof(double), nk, cudaMemcpyHostToDevice, R.stream));
  CLV_CUDA(cudaMemcpy2DAsync(d(nj + 1, 1), dp, h(nj + 1, 1), sp, (size_t)(jmax - nj) * sizeof(double), nk,
                             cudaMemcpyHostToDevice, R.stream));
  R.h2d += (long long)(sp * (2 + kmax - nk) + (size_t)(2 + jmax - nj) * sizeof(double) * nk);
}

void download(const Entry& e, double* host) {
  if (is_2d(e.kind)) {
    const int pitch = pitch_for(e.nx);
    const size_t w = (size_t)host_row(e.kind, e.nx) * sizeof(double);
    const size_t h = host_rows(e.kind, e.ny);
    CLV_CUDA(cudaMemcpy2DAsync(host, w, e.d + (XOFF - 1), (size_t)pitch * sizeof(double), w, h,
                               cudaMemcpyDeviceToHost, R.stream));
    R.d2h += (long long)(w * h);
  } else {
    const size_t n = len_1d(e.kind, e.nx, e.ny) * sizeof(double);
    CLV_CUDA(cudaMemcpyAsync(host, e.d, n, cudaMemcpyDeviceToHost, R.stream));
    R.d2h += (long long)n;
  }
}

// Device allocations are pooled by size: a mirror that is forgotten (clover_b200_forget_, a host address re-used with
// another shape) parks its buffers here and the next mirror of that size takes them -- no cudaMalloc / cudaFree on
// the path of a run that re-creates its arrays, and tensor maps encoded for a parked buffer stay valid (they depend
// on address and shape only).  clover_b200_invalidate_ / finalize_ really free the pool.
double* alloc_zeroed(size_t doubles) {
  double* p = nullptr;
  auto it = R.pool.find(doubles);
  if (it != R.pool.end()) {
    p = it->second;
    R.pool.erase(it);
  } else {
    CLV_CUDA(cudaMalloc(&p, doubles * sizeof(double)));
  }
  CLV_CUDA(cudaMemsetAsync(p, 0, doubles * sizeof(double), R.stream));
  return p;
}
void park(double* p, size_t doubles) {
  if (p) R.pool.emplace(doubles, p);
}
void free_pool() {
  for (auto& kv : R.pool) CLV_CUDA(cudaFree(kv.second));
  R.pool.clear();
}

size_t doubles_for(Kind kind, int nx, int ny) {
  // uniform 2-D shape: (ny+5) rows of `pitch`, plus a row of slack so that clamped halo threads
  // of the sweep kernels can never leave the allocation
  if (is_2d(kind)) return (size_t)pitch_for(nx) * (size_t)(ny + 6);
  return len_1d(kind, nx, ny) + 8;
}

// The chunk registered for exchange / sync_to_host (one chunk per process == per GPU)
struct Chunk {
  bool set = false;
  int nx = 0, ny = 0;
  int neighbours[4] = {-1, -1, -1, -1};
  double* field[15] = {};
} C;

}  // namespace

void ensure_init() {
  if (R.ready) return;
  int dev_id = 0;
  if (const char* s = getenv("CLOVER_B200_DEVICE")) dev_id = atoi(s);
  clover_b200_init_(&dev_id);
}

cudaStream_t stream() { return R.cur ? R.cur : R.stream; }

// Overlap of a halo exchange with the work that does not depend on it (fuse.cu: the viscosity exchange runs next
// to the dt reduction, the host's dt read and the PdV predictor).  side_begin(): what is issued from now on goes to
// the side stream, ordered after everything already on the main stream; side_end(): back to the main stream;
// join_side(): the main stream waits for the side work -- called before anything that could touch its data.
bool overlap_enabled() { return R.overlap && !R.profiling && R.resident; }
void side_begin() {
  CLV_CUDA(cudaEventRecord(R.ev_fork, R.stream));
  CLV_CUDA(cudaStreamWaitEvent(R.side, R.ev_fork, 0));
  R.cur = R.side;
}
void side_end() {
  R.cur = nullptr;
  CLV_CUDA(cudaEventRecord(R.ev_join, R.side));
  R.side_pending = true;
}
void join_side() {
  if (!R.side_pending) return;
  CLV_CUDA(cudaStreamWaitEvent(R.stream, R.ev_join, 0));
  R.side_pending = false;
}

bool is_resident() { return R.resident; }

Grid grid_of(const int* xmin, const int* xmax, const int* ymin, const int* ymax) {
  flush_deferred();
  join_side();
  return grid_of_noflush(xmin, xmax, ymin, ymax);
}

bool fusion_enabled() { return R.fuse; }
bool tma_enabled() { return R.tma; }
int sm_count() { return R.sms; }

void submit(Op&& op) {
  if (!R.resident) {
    op.run();
    finish();
    return;
  }
  R.queue.push_back(std::move(op));
  if (R.queue.size() >= 1024) flush_deferred();
}

bool deferred_queue_empty() { return R.queue.empty(); }

void flush_deferred() {
  if (R.draining) return;
  // a held-back viscosity halo update (fuse.cu) waits for the PdV predictor pattern only; anything else first
  // (another call, or a host-visible point with nothing recorded) gets it issued now
  if (pending_halo_exists() && (R.queue.empty() || R.queue[0].kind != OP_PDV_PREDICT)) {
    R.draining = true;
    run_pending_halo();
    R.draining = false;
  }
  if (R.queue.empty()) return;
  R.draining = true;
  std::vector<Op> q;
  q.swap(R.queue);
  for (size_t i = 0; i < q.size();) {
    // the PdV predictor pattern (fuse.cu) joins the side stream itself, after its compute launch
    if (q[i].kind != OP_PDV_PREDICT) join_side();
    size_t used = R.fuse ? fuse_at(q.data(), q.size(), i) : 0;
    if (used == 0) {
      join_side();
      if (pending_halo_exists()) run_pending_halo();  // (the predictor pattern did not match: nothing merges it)
      q[i].run();
      used = 1;
    }
    i += used;
  }
  R.draining = false;
}

Grid grid_of_noflush(const int* xmin, const int* xmax, const int* ymin, const int* ymax) {
  ensure_init();
  if (*xmin != 1 || *ymin != 1)
    fatal("x_min/y_min must be 1 (start.f90:77-80 always passes 1), got %d/%d", *xmin, *ymin);
  if (*xmax < 1 || *ymax < 1) fatal("empty chunk %d x %d", *xmax, *ymax);
  Grid g;
  g.nx = *xmax;
  g.ny = *ymax;
  g.pitch = pitch_for(g.nx);
  return g;
}

static void materialize(Entry& e);
static void materialize_dependents(const double* host);
static void materialize_all();
static void drop_lazy(Entry& e) {
  if (e.lazy_src) {
    e.lazy_src = nullptr;
    e.lazy_src2 = nullptr;
    R.lazy_pending--;
  }
}

static Entry& lookup(const Grid& g, const double* host, Kind kind, bool* fresh) {
  if (!host) fatal("null array pointer passed to a kernel entry point");
  auto it = R.arrays.find(host);
  if (it != R.arrays.end()) {
    Entry& e = it->second;
    if (e.kind == kind && e.nx == g.nx && e.ny == g.ny) {
      *fresh = false;
      return e;
    }
    // same host address re-used with another shape (e.g. a work array): re-create the mirror
    materialize_dependents(host);
    drop_lazy(e);
    park(e.d, e.doubles);
    park(e.alt, e.doubles);
    R.arrays.erase(it);
  }
  Entry e;
  e.kind = kind;
  e.nx = g.nx;
  e.ny = g.ny;
  e.doubles = doubles_for(kind, g.nx, g.ny);
  e.d = alloc_zeroed(e.doubles);
  *fresh = true;
  return R.arrays.emplace(host, e).first->second;
}

// ---- lazy copies ----------------------------------------------------------------------------------
void launch_copy_range(const Grid& g, const double* src, double* dst, Kind kind);  // lagrange.cu
void launch_soundspeed(const Grid& g, const double* density, const double* energy, double* soundspeed);  // lagrange.cu

static void materialize(Entry& e) {
  if (!e.lazy_src) return;
  auto it = R.arrays.find(e.lazy_src);
  if (it == R.arrays.end()) fatal("lazy copy source vanished");
  Grid g{e.nx, e.ny, pitch_for(e.nx)};
  if (e.lazy_src2) {
    auto i2 = R.arrays.find(e.lazy_src2);
    if (i2 == R.arrays.end()) fatal("lazy soundspeed source vanished");
    launch_soundspeed(g, it->second.d, i2->second.d, e.d);
  } else {
    launch_copy_range(g, it->second.d, e.d, e.kind);
  }
  e.lazy_src = nullptr;
  e.lazy_src2 = nullptr;
  R.lazy_pending--;
}
// `host` is about to be modified: perform every pending copy / evaluation that reads from it
static void materialize_dependents(const double* host) {
  if (R.lazy_pending == 0) return;
  for (auto& kv : R.arrays)
    if (kv.second.lazy_src == host || kv.second.lazy_src2 == host) materialize(kv.second);
}
static void materialize_all() {
  if (R.lazy_pending == 0) return;
  for (auto& kv : R.arrays) materialize(kv.second);
}

double* dev(const Grid& g, const double* host, Kind kind, int access) {
  bool fresh = false;
  Entry& e = lookup(g, host, kind, &fresh);
  if (R.lazy_pending) {
    if (e.lazy_src) {
      if (access == OUT_FULL) {
        drop_lazy(e);
      } else {
        materialize(e);
      }
    }
    if (access & OUT) materialize_dependents(host);
  }
  // Resident mode: the host copy is authoritative only the first time the address is seen.
  // Copy-in/out mode: it is authoritative on every call (outputs too: a kernel writes only its
  // loop range, the rest of the array must survive the round trip).
  if (fresh && R.resident && access == OUT_FULL && is_2d(kind)) upload_ring(e, host);
  else if (fresh || !R.resident) upload(e, host);
  if (!R.resident && (access & (OUT | HALO))) R.pending_out.push_back(host);
  return e.d;
}

void lazy_copy(const Grid& g, const double* dst_host, const double* src_host, Kind kind) {
  dev(g, src_host, kind, IN);        // exists, uploaded, itself materialised
  dev(g, dst_host, kind, OUT_FULL);  // exists; whatever was pending for dst is superseded; dependents of dst done
  Entry& d = R.arrays.find(dst_host)->second;
  d.lazy_src = src_host;
  R.lazy_pending++;
}

// soundspeed := c(density, energy) on the interior (ideal_gas_kernel_c.c:48-59), recorded instead of evaluated: the
// fused timestep launch consumes the sound speed on chip, and in the reference's call order the array is overwritten
// (PdV predictor's ideal_gas) or dead before anything reads it.  Any read (calc_dt on its own, a download), or a
// write to density / energy, evaluates it first; an overwrite drops it.
void lazy_soundspeed(const Grid& g, const double* ss_host, const double* d_host, const double* e_host) {
  dev(g, d_host, CELL, IN);
  dev(g, e_host, CELL, IN);
  dev(g, ss_host, CELL, OUT_FULL);
  Entry& s = R.arrays.find(ss_host)->second;
  s.lazy_src = d_host;
  s.lazy_src2 = e_host;
  R.lazy_pending++;
}

void swap_buffers(const double* host_a, const double* host_b) {
  auto a = R.arrays.find(host_a), b = R.arrays.find(host_b);
  if (a == R.arrays.end() || b == R.arrays.end()) fatal("swap_buffers on an unknown array");
  Entry &x = a->second, &y = b->second;
  if (x.kind != y.kind || x.nx != y.nx || x.ny != y.ny) fatal("swap_buffers: shape mismatch");
  if (x.lazy_src || y.lazy_src) fatal("swap_buffers with a lazy copy pending");
  // pending copies that READ one of the two arrays must see its contents from before the swap
  materialize_dependents(host_a);
  materialize_dependents(host_b);
  std::swap(x.d, y.d);
}

double* dev_alt(const Grid& g, const double* host, Kind kind) {
  bool fresh = false;
  Entry& e = lookup(g, host, kind, &fresh);
  if (fresh) upload(e, host);
  if (!e.alt) e.alt = alloc_zeroed(e.doubles);
  return e.alt;
}

void swap_alt(const double* host) {
  auto it = R.arrays.find(host);
  if (it == R.arrays.end() || !it->second.alt) fatal("swap_alt on an array without alt buffer");
  std::swap(it->second.d, it->second.alt);
}

double* dev_buffer(const double* host, size_t need, int access, size_t lo, size_t hi) {
  ensure_init();
  if (!host) fatal("null message buffer");
  Buffer& b = R.buffers[host];
  if (b.doubles < need) {
    size_t n = need + need / 2 + 64;
    double* p = alloc_zeroed(n);
    if (b.d) {
      CLV_CUDA(cudaMemcpyAsync(p, b.d, b.doubles * sizeof(double), cudaMemcpyDeviceToDevice, R.stream));
      CLV_CUDA(cudaStreamSynchronize(R.stream));
      CLV_CUDA(cudaFree(b.d));
    }
    b.d = p;
    b.doubles = n;
  }
  if ((access & IN) && hi > lo) {
    CLV_CUDA(cudaMemcpyAsync(b.d + lo, host + lo, (hi - lo) * sizeof(double), cudaMemcpyHostToDevice, R.stream));
    R.h2d += (long long)((hi - lo) * sizeof(double));
  }
  return b.d;
}

void finish() {
  if (R.resident) return;
  for (const void* h : R.pending_out) {
    auto it = R.arrays.find(h);
    if (it != R.arrays.end()) download(it->second, (double*)h);
  }
  R.pending_out.clear();
  CLV_CUDA(cudaStreamSynchronize(R.stream));
}

LaunchScope::LaunchScope(const char* n) : name(n), trace(nullptr) {
  if (R.profiling) CLV_CUDA(cudaEventRecord(R.ev0, R.stream));
  if (R.trace_on && R.trace_names.size() < TRACE_CAP) {
    trace = R.d_trace + 8 * R.trace_names.size();
    R.trace_names.push_back(n);
  }
  R.cur_trace = trace;
}
unsigned long long* current_trace() { return R.cur_trace; }

LaunchScope::~LaunchScope() {
  R.launches++;
  R.cur_trace = nullptr;
  R.halo_noted = false;  // halo.cu re-notes after its scope closes; any other launch ends the "just launched" state
  R.ring_swap_noted = false;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fatal("launch of %s failed: %s", name, cudaGetErrorString(e));
  if (R.profiling) {
    CLV_CUDA(cudaEventRecord(R.ev1, R.stream));
    CLV_CUDA(cudaEventSynchronize(R.ev1));
    float ms = 0;
    CLV_CUDA(cudaEventElapsedTime(&ms, R.ev0, R.ev1));
    Prof& p = R.prof[name];
    p.ms += ms;
    p.calls++;
  }
}

double* host_scalars() { return R.h_scalars; }
double* device_error_record() { return R.h_scalars + 48; }
unsigned long long spin_timeout_ns() { return R.spin_timeout_ns; }

// A kernel that gave up waiting for another GPU (or for its own grid) left a record in pinned host memory before it
// trapped; the CUDA error that follows the trap brings us here.
void report_device_error() {
  if (!R.h_scalars) return;
  const volatile double* e = R.h_scalars + 48;
  const int code = (int)e[0];
  if (code == 0) return;
  static const char* what[] = {"", "halo exchange: a neighbour's strips never arrived", "halo exchange: grid barrier never completed",
                               "all-reduce: a rank's contribution never arrived"};
  fprintf(stderr,
          "libclover_b200: device-side time-out (%.1f s) on rank %d: %s (peer/slot %d, waiting for sequence number %.0f, "
          "last seen %.0f).  A peer process has died or fallen out of step; aborting instead of hanging the GPU.\n",
          (double)R.spin_timeout_ns * 1e-9, (int)e[1], (code >= 1 && code <= 3) ? what[code] : "unknown wait", (int)e[2], e[3], e[4]);
}

ReduceTail next_reduce_tail(int base) {
  ReduceTail t;
  t.seq = (double)(++R.scalar_seq);
  t.all = nullptr;
  t.nranks = 1;
  t.rank = 0;
  t.ar_seq = 0;
  t.timeout_ns = R.spin_timeout_ns;
  t.err = R.h_scalars + 48;
  (void)base;
  fill_reduce_tail_ranks(t);
  return t;
}

void wait_scalars(int base, double seq) {
  const volatile double* flag = R.h_scalars + base + 7;
  unsigned long long spins = 0;
  struct timeval t0;
  gettimeofday(&t0, nullptr);
  while (*flag != seq) {
    if ((++spins & 0xfffff) == 0) {  // ~every millisecond: a stream query takes microseconds and must not sit on the dt path
      const cudaError_t q = cudaStreamQuery(R.stream);
      if (q != cudaSuccess && q != cudaErrorNotReady) CLV_CUDA(q);
      if (q == cudaSuccess && *flag != seq) {  // the stream has drained and the kernel never wrote: cannot happen
        if (*flag != seq) fatal("reduction result %d never arrived (stream idle)", base);
      }
      struct timeval t1;
      gettimeofday(&t1, nullptr);
      if (t1.tv_sec - t0.tv_sec > 60) {
        report_device_error();
        fatal("reduction result %d not delivered after 60 s", base);
      }
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  __sync_synchronize();
}
unsigned int* ticket() { return R.d_ticket; }
double* partials(size_t doubles) {
  if (R.partials_doubles < doubles) {
    if (R.d_partials) {
      CLV_CUDA(cudaStreamSynchronize(R.stream));
      CLV_CUDA(cudaFree(R.d_partials));
    }
    R.partials_doubles = doubles + 1024;
    CLV_CUDA(cudaMalloc(&R.d_partials, R.partials_doubles * sizeof(double)));
  }
  return R.d_partials;
}

// ---- TMA tensor maps (tma.cuh) ------------------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver-API entry point; it is fetched through the runtime so that the library
// does not link libcuda.  A map depends only on (address, pitch, rows, box), so entries stay valid for as long
// as an allocation of that shape lives at that address; the cache is dropped whenever mirrors are freed.
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
struct MapKey {
  const void* p;
  int pitch, rows, bw, bh;
  bool operator<(const MapKey& o) const {
    return std::tie(p, pitch, rows, bw, bh) < std::tie(o.p, o.pitch, o.rows, o.bw, o.bh);
  }
};
std::map<MapKey, CUtensorMap*> g_maps;
void drop_tensor_maps() {
  for (auto& kv : g_maps) free(kv.second);
  g_maps.clear();
}
}  // namespace

// L2 fetch granularity of the TMA loads (0 none, 1 64 B, 2 128 B, 3 256 B); $CLOVER_B200_L2PROMO overrides for A/B runs
static int l2_promotion() {
  static int v = -1;
  if (v < 0) v = getenv("CLOVER_B200_L2PROMO") ? atoi(getenv("CLOVER_B200_L2PROMO")) : 3;
  return v;
}

const CUtensorMap* tensor_map_for(const Grid& g, const double* dev_ptr, int box_w, int box_h) {
  const MapKey key{dev_ptr, g.pitch, g.ny + 6, box_w, box_h};
  auto it = g_maps.find(key);
  if (it != g_maps.end()) return it->second;
  if (!g_encode_tiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CLV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) fatal("cuTensorMapEncodeTiled is not available in this driver");
    g_encode_tiled = (EncodeTiledFn)fn;
  }
  CUtensorMap* m = nullptr;
  if (posix_memalign((void**)&m, 64, sizeof(CUtensorMap)) != 0) fatal("out of host memory");
  const cuuint64_t dims[2] = {(cuuint64_t)g.pitch, (cuuint64_t)(g.ny + 6)};
  const cuuint64_t strides[1] = {(cuuint64_t)g.pitch * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)dev_ptr, dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    (CUtensorMapL2promotion)l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fatal("cuTensorMapEncodeTiled failed (%d) for pitch %d rows %d box %dx%d", (int)r, g.pitch, g.ny + 6, box_w, box_h);
  g_maps[key] = m;
  return m;
}

namespace {
std::map<std::tuple<int, int, int, int, int, int, int, int, int, int>, TileOrder> g_split_orders;
}
TileOrder tile_order_split(int ntx, int nty, int tw, int th, int lo_x, int hi_x, int lo_y, int hi_y, int nx, int ny) {
  const auto key = std::make_tuple(ntx, nty, tw, th, lo_x, hi_x, lo_y, hi_y, nx, ny);
  auto it = g_split_orders.find(key);
  if (it != g_split_orders.end()) return it->second;
  std::vector<TileXY> h;
  const int n_interior = build_tile_order(ntx, nty, tw, th, lo_x, hi_x, lo_y, hi_y, nx, ny, h);  // tile_order.h
  g_split_orders[key].n_interior = n_interior;
  if ((int)h.size() != ntx * nty) fatal("tile_order_split: %zu entries for %d x %d tiles", h.size(), ntx, nty);
  static_assert(sizeof(TileXY) == sizeof(int2), "TileXY is uploaded as int2");
  int2* d = nullptr;
  CLV_CUDA(cudaMalloc(&d, h.size() * sizeof(int2)));
  CLV_CUDA(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(int2), cudaMemcpyHostToDevice, R.stream));
  CLV_CUDA(cudaStreamSynchronize(R.stream));
  TileOrder& o = g_split_orders[key];
  o.table = d;
  o.ntiles = ntx * nty;
  return o;
}

// Ticket counters of the dynamic tile queues (tma.cuh): a ring of zero-initialised {tickets, exits} pairs, one pair
// per launch; the launch itself zeroes its pair when its last CTA leaves.  The ring is far longer than the number of
// grids that can be in flight at once (a PDL chain of tiny kernels can have many resident at the same time).
constexpr unsigned int TICKET_RING = 2048;
Tickets next_tickets() {
  static int dynamic = -1;
  if (dynamic < 0) {
    const char* e = getenv("CLOVER_B200_QUEUE");
    dynamic = (e && strcmp(e, "static") == 0) ? 0 : 1;
  }
  if (!dynamic) return Tickets{nullptr};
  if (!R.d_tickets) {
    CLV_CUDA(cudaMalloc(&R.d_tickets, TICKET_RING * 2 * sizeof(unsigned int)));
    CLV_CUDA(cudaMemsetAsync(R.d_tickets, 0, TICKET_RING * 2 * sizeof(unsigned int), R.stream));
  }
  Tickets t;
  t.ctr = R.d_tickets + 2 * (R.ticket_turn++ % TICKET_RING);
  return t;
}

// ---- programmatic dependent launch bookkeeping (common.cuh) ----------------------------------------------------
bool pdl_enabled() { return R.pdl && !R.profiling; }
void note_halo_launch() { R.halo_noted = true; }
bool halo_just_launched() { return R.halo_noted; }
// reset_field's ring swap touches halo rings only and triggers its dependents after its own wait: the halo kernel that
// follows it may trigger BEFORE its wait (its dependents' interior tiles then run next to the ring swap as well)
void note_ring_swap_launch() { R.ring_swap_noted = true; }
bool ring_swap_just_launched() { return R.ring_swap_noted && pdl_enabled(); }
int dep_start_for(const TileOrder& o) { return (halo_just_launched() && pdl_enabled() && R.split) ? o.n_interior : 0; }

// used by halo.cu
bool chunk_registered() { return C.set; }
int chunk_nx() { return C.nx; }
int chunk_ny() { return C.ny; }
const int* chunk_neighbours() { return C.neighbours; }
double* chunk_field_host(int f) { return C.field[f]; }
void count_copy(long long h2d, long long d2h) { R.h2d += h2d; R.d2h += d2h; }

}  // namespace clv

using namespace clv;

extern "C" {

void clover_b200_init_(int* device) {
  if (R.ready) return;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    fatal("no CUDA device available (%s); this library has no CPU fallback",
          e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  R.device = device ? *device : 0;
  if (R.device < 0 || R.device >= n) fatal("device %d out of range (%d devices)", R.device, n);
  CLV_CUDA(cudaSetDevice(R.device));
  cudaDeviceProp p;
  CLV_CUDA(cudaGetDeviceProperties(&p, R.device));
  if (p.major != 10)
    fatal("device %d is sm_%d%d; this library is built for sm_100a (B200) only", R.device, p.major, p.minor);
  CLV_CUDA(cudaStreamCreateWithFlags(&R.stream, cudaStreamNonBlocking));
  CLV_CUDA(cudaStreamCreateWithFlags(&R.side, cudaStreamNonBlocking));
  CLV_CUDA(cudaEventCreateWithFlags(&R.ev_fork, cudaEventDisableTiming));
  CLV_CUDA(cudaEventCreateWithFlags(&R.ev_join, cudaEventDisableTiming));
  if (const char* s = getenv("CLOVER_B200_OVERLAP")) R.overlap = (atoi(s) != 0);
  CLV_CUDA(cudaHostAlloc(&R.h_scalars, 64 * sizeof(double), cudaHostAllocMapped));
  CLV_CUDA(cudaMalloc(&R.d_ticket, 16 * sizeof(unsigned int)));
  CLV_CUDA(cudaMemset(R.d_ticket, 0, 16 * sizeof(unsigned int)));
  R.ticket_turn = 0;
  CLV_CUDA(cudaEventCreate(&R.ev0));
  CLV_CUDA(cudaEventCreate(&R.ev1));
  if (const char* s = getenv("CLOVER_B200_PDL")) R.pdl = (atoi(s) != 0);
  if (const char* s = getenv("CLOVER_B200_SPIN_TIMEOUT_MS")) R.spin_timeout_ns = (unsigned long long)atoll(s) * 1000000ull;
  memset(R.h_scalars, 0, 64 * sizeof(double));
  if (const char* s = getenv("CLOVER_B200_SPLIT")) R.split = (atoi(s) != 0);
  if (const char* s = getenv("CLOVER_B200_FUSE")) R.fuse = (atoi(s) != 0);
  if (const char* s = getenv("CLOVER_B200_TMA")) R.tma = (atoi(s) != 0);  // A/B switch for profiling
  R.sms = p.multiProcessorCount;
  R.ready = true;
}

void clover_b200_comm_finalize_internal();

void clover_b200_finalize_(void) {
  if (!R.ready) return;
  flush_deferred();
  join_side();
  CLV_CUDA(cudaStreamSynchronize(R.stream));
  clover_b200_comm_finalize_internal();
  clover_b200_invalidate_();
  if (R.d_partials) CLV_CUDA(cudaFree(R.d_partials));
  R.d_partials = nullptr;
  R.partials_doubles = 0;
  CLV_CUDA(cudaFree(R.d_ticket));
  if (R.d_tickets) CLV_CUDA(cudaFree(R.d_tickets));
  R.d_tickets = nullptr;
  CLV_CUDA(cudaFreeHost(R.h_scalars));
  R.h_scalars = nullptr;
  CLV_CUDA(cudaEventDestroy(R.ev0));
  CLV_CUDA(cudaEventDestroy(R.ev1));
  CLV_CUDA(cudaStreamDestroy(R.stream));
  CLV_CUDA(cudaStreamDestroy(R.side));
  CLV_CUDA(cudaEventDestroy(R.ev_fork));
  CLV_CUDA(cudaEventDestroy(R.ev_join));
  R.side_pending = false;
  R.ready = false;
  C = Chunk();
}

void clover_b200_set_resident_(int* on) {
  ensure_init();
  flush_deferred();
  join_side();
  materialize_all();
  CLV_CUDA(cudaStreamSynchronize(R.stream));
  R.resident = (*on != 0);
}

void clover_b200_invalidate_(void) {
  if (!R.ready) return;
  flush_deferred();
  join_side();
  CLV_CUDA(cudaStreamSynchronize(R.stream));
  for (auto& kv : R.arrays) {
    CLV_CUDA(cudaFree(kv.second.d));
    if (kv.second.alt) CLV_CUDA(cudaFree(kv.second.alt));
  }
  R.arrays.clear();
  free_pool();
  drop_tensor_maps();
  R.lazy_pending = 0;
  for (auto& kv : R.buffers) CLV_CUDA(cudaFree(kv.second.d));
  R.buffers.clear();
  R.pending_out.clear();
}

void clover_b200_forget_(double* host) {
  if (!R.ready) return;
  flush_deferred();
  join_side();
  auto it = R.arrays.find(host);
  if (it != R.arrays.end()) {
    materialize_dependents(host);
    drop_lazy(it->second);
    park(it->second.d, it->second.doubles);  // stream order keeps a later user of the buffer behind its last use here
    park(it->second.alt, it->second.doubles);
    R.arrays.erase(it);
  }
  auto ib = R.buffers.find(host);
  if (ib != R.buffers.end()) {
    CLV_CUDA(cudaStreamSynchronize(R.stream));
    CLV_CUDA(cudaFree(ib->second.d));
    R.buffers.erase(ib);
  }
}

void clover_b200_upload_(double* host) {
  ensure_init();
  flush_deferred();
  join_side();
  auto it = R.arrays.find(host);
  if (it == R.arrays.end()) return;  // never seen: the first use uploads it anyway
  materialize_dependents(host);
  drop_lazy(it->second);
  upload(it->second, host);
}

void clover_b200_download_(double* host) {
  ensure_init();
  flush_deferred();
  join_side();
  auto it = R.arrays.find(host);
  if (it == R.arrays.end()) return;  // never seen by a kernel: the host copy is the only copy, nothing to bring back
  materialize(it->second);
  download(it->second, host);
  CLV_CUDA(cudaStreamSynchronize(R.stream));
}

void clover_b200_sync_to_host_(int* fields) {
  ensure_init();
  flush_deferred();
  join_side();
  if (!C.set) fatal("sync_to_host before register_chunk");
  for (int f = 0; f < 15; ++f) {
    if (fields && fields[f] != 1) continue;
    auto it = R.arrays.find(C.field[f]);
    if (it != R.arrays.end()) {
      materialize(it->second);
      download(it->second, C.field[f]);
    }
  }
  CLV_CUDA(cudaStreamSynchronize(R.stream));
}

void clover_b200_device_synchronize_(void) {
  ensure_init();
  flush_deferred();
  join_side();
  CLV_CUDA(cudaStreamSynchronize(R.stream));
}

void clover_b200_register_chunk_(int* xmin, int* xmax, int* ymin, int* ymax, int* nb, double* density0,
                                 double* density1, double* energy0, double* energy1, double* pressure,
                                 double* viscosity, double* soundspeed, double* xvel0, double* xvel1,
                                 double* yvel0, double* yvel1, double* vol_flux_x, double* vol_flux_y,
                                 double* mass_flux_x, double* mass_flux_y) {
  Grid g = grid_of(xmin, xmax, ymin, ymax);
  C.set = true;
  C.nx = g.nx;
  C.ny = g.ny;
  for (int i = 0; i < 4; ++i) C.neighbours[i] = nb[i];
  // field ids of data.f90:51-66, zero-based
  double* f[15] = {density0, density1, energy0, energy1, pressure, viscosity, soundspeed, xvel0,
                   xvel1, yvel0, yvel1, vol_flux_x, vol_flux_y, mass_flux_x, mass_flux_y};
  for (int i = 0; i < 15; ++i) C.field[i] = f[i];
}

void clover_b200_launch_count_(long long* n) {
  flush_deferred();
  *n = R.launches;
}

void clover_b200_profile_(int* on) {
  ensure_init();
  flush_deferred();
  R.profiling = (*on != 0);
}

void clover_b200_set_fusion_(int* on) {
  ensure_init();
  flush_deferred();
  R.fuse = (*on != 0);
}

void clover_b200_set_tma_(int* on) {
  ensure_init();
  flush_deferred();
  R.tma = (*on != 0);
}

void clover_b200_profile_reset_(void) { R.prof.clear(); }

// In-situ timeline: with *on != 0 every launch from now on gets a slot of four %globaltimer stamps (common.cuh);
// trace_dump_ writes "index,name,start_ns,end_ns,wait_begin_ns,wait_end_ns,packed_ns,unpacked_ns,barrier_ns" (relative to the first start; empty
// fields where a kernel has no such stamp) and clears the record.  Unlike the event profile this does not serialise
// the launches: it shows the step as it really runs (overlap of the halo exchange with interior tiles, gaps).
void clover_b200_trace_(int* on) {
  ensure_init();
  flush_deferred();
  join_side();
  CLV_CUDA(cudaStreamSynchronize(R.stream));
  if (*on && !R.d_trace) CLV_CUDA(cudaMalloc(&R.d_trace, TRACE_CAP * 8 * sizeof(unsigned long long)));
  if (*on) {
    std::vector<unsigned long long> init(TRACE_CAP * 8);
    for (size_t i = 0; i < TRACE_CAP; ++i) {
      for (int k = 0; k < 8; ++k) init[8 * i + k] = 0;  // stamps folded with max
      init[8 * i] = ~0ull;                               // [0], [2]: folded with min
      init[8 * i + 2] = ~0ull;
    }
    CLV_CUDA(cudaMemcpy(R.d_trace, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    R.trace_names.clear();
  }
  R.trace_on = (*on != 0);
}
void clover_b200_trace_dump_(const char* path) {
  ensure_init();
  flush_deferred();
  join_side();
  CLV_CUDA(cudaStreamSynchronize(R.stream));
  const size_t n = R.trace_names.size();
  std::vector<unsigned long long> h(8 * (n ? n : 1));
  if (n) CLV_CUDA(cudaMemcpy(h.data(), R.d_trace, 8 * n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  FILE* f = fopen(path, "w");
  if (!f) fatal("trace_dump: cannot write %s", path);
  unsigned long long t0 = ~0ull;
  for (size_t i = 0; i < n; ++i)
    if (h[8 * i] < t0) t0 = h[8 * i];
  fprintf(f, "index,name,start_ns,end_ns,wait_begin_ns,wait_end_ns,packed_ns,unpacked_ns,barrier_ns\n");
  for (size_t i = 0; i < n; ++i) {
    fprintf(f, "%zu,%s,", i, R.trace_names[i]);
    if (h[8 * i] != ~0ull) fprintf(f, "%llu", h[8 * i] - t0);
    fprintf(f, ",");
    if (h[8 * i + 1] != 0) fprintf(f, "%llu", h[8 * i + 1] - t0);
    fprintf(f, ",");
    if (h[8 * i + 2] != ~0ull) fprintf(f, "%llu", h[8 * i + 2] - t0);
    for (int k = 3; k < 7; ++k) {
      fprintf(f, ",");
      if (h[8 * i + k] != 0) fprintf(f, "%llu", h[8 * i + k] - t0);
    }
    fprintf(f, "\n");
  }
  fclose(f);
  R.trace_names.clear();
  R.trace_on = false;
}

void clover_b200_profile_get_(int* max, char* names32, double* total_ms, long long* calls, int* n) {
  int i = 0;
  for (auto& kv : R.prof) {
    if (i >= *max) break;
    memset(names32 + 32 * i, 0, 32);
    strncpy(names32 + 32 * i, kv.first.c_str(), 31);
    total_ms[i] = kv.second.ms;
    calls[i] = kv.second.calls;
    ++i;
  }
  *n = i;
}

void clover_b200_copy_bytes_(long long* h2d, long long* d2h) {
  *h2d = R.h2d;
  *d2h = R.d2h;
}

void timer_c_(double* elapsed_time) {
  struct timeval t;
  gettimeofday(&t, nullptr);
  *elapsed_time = t.tv_sec + t.tv_usec * 1.0E-6;
}

}  // extern "C"

// ---- device-side timing and host pinning for bench.py ---------------------------------------------
namespace {
cudaEvent_t g_events[8] = {};
}
extern "C" {
// Record event `slot` (0..7) on the library's stream.
void clover_b200_event_record_(int* slot) {
  clv::ensure_init();
  clv::flush_deferred();
  clv::join_side();
  if (*slot < 0 || *slot >= 8) clv::fatal("event slot %d", *slot);
  if (!g_events[*slot]) CLV_CUDA(cudaEventCreate(&g_events[*slot]));
  CLV_CUDA(cudaEventRecord(g_events[*slot], clv::stream()));
}
// Milliseconds between two recorded events (waits for the later one).
void clover_b200_event_elapsed_ms_(int* a, int* b, double* ms) {
  CLV_CUDA(cudaEventSynchronize(g_events[*b]));
  float f = 0;
  CLV_CUDA(cudaEventElapsedTime(&f, g_events[*a], g_events[*b]));
  *ms = f;
}
// Page-lock a host array so that uploads/downloads run at full PCIe rate (optional).
void clover_b200_pin_(double* host, long long* bytes) {
  clv::ensure_init();
  cudaError_t e = cudaHostRegister(host, (size_t)*bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess && e != cudaErrorHostMemoryAlreadyRegistered)
    clv::fatal("cudaHostRegister(%lld bytes): %s", *bytes, cudaGetErrorString(e));
  (void)cudaGetLastError();
}
void clover_b200_unpin_(double* host) {
  (void)cudaHostUnregister(host);
  (void)cudaGetLastError();
}
}
