// Host driver: a C++ restatement of the CALL SEQUENCE of CloverLeaf_ref's
// Fortran driver (L0/L1 of SURVEY.md), written against the reference's own
// `*_kernel_c_` C entry points (lowercase + trailing underscore, every argument
// by reference).  The same driver therefore runs any backend that exports that
// ABI: the CUDA library (libclover_b200.so, the product) or -- in tests and the
// CPU-baseline leg of bench.py only -- the checkers built under oracle/.  The
// driver never picks a backend itself: the caller hands it one shared library.
//
// No Fortran compiler / MPI exists in this image, so this file stands in for
// the Fortran driver that "stays" in the north-star design; it contains NO
// kernel arithmetic.  What each routine follows (all under
// /root/reference/CloverLeaf_ref/):
//   parse_deck        read_input.f90:38-71,128-191,274-281 ; parse.f90:160-184
//   decompose         clover.f90:108-206
//   start             start.f90:48-143 ; build_field.f90:33-181 ; clover.f90:329-342
//   initialise_chunk  initialise_chunk.f90:33-38
//   generate_chunk    generate_chunk.f90:36-47
//   hydro_step        hydro.f90:46-84
//   timestep          timestep.f90:56-117 ; calc_dt.f90
//   pdv               PdV.f90:46-138
//   advection         advection.f90:43-110 ; advec_cell_driver.f90 ; advec_mom_driver.f90:83-131
//   field_summary     field_summary.f90:53-129
//   update_halo       update_halo.f90:39-113
//   exchange          clover.f90:348-500 (+ the field->data-type table of :690-880)
//
// Multi-chunk without MPI: with comm_mode==0 ALL chunks of the decomposition
// live in this process and `exchange` copies each chunk's send buffer into the
// neighbour's receive buffer (what MPI_ISEND/IRECV do in the reference).  With
// comm_mode==1 this process owns exactly one chunk (rank+1) and the exchange /
// reductions are delegated to the backend's `clover_b200_*` extension entry
// points (device pack + NCCL), see include/clover_b200.h.
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

// data.f90:30-31 -- default-REAL literals assigned to REAL(KIND=8) parameters.
const double g_small = (double)1.0e-16f;
const double g_big = (double)1.0e+21f;
const int g_ibig = 640000;
enum { G_RECT = 1, G_CIRC = 2, G_POINT = 3 };                   // data.f90
enum { CELL_DATA = 1, VERTEX_DATA = 2, X_FACE_DATA = 3, Y_FACE_DATA = 4 };
enum { LEFT = 0, RIGHT = 1, BOTTOM = 2, TOP = 3 };              // CHUNK_LEFT-1 ...
enum {
  FIELD_DENSITY0 = 1, FIELD_DENSITY1, FIELD_ENERGY0, FIELD_ENERGY1, FIELD_PRESSURE,
  FIELD_VISCOSITY, FIELD_SOUNDSPEED, FIELD_XVEL0, FIELD_XVEL1, FIELD_YVEL0, FIELD_YVEL1,
  FIELD_VOL_FLUX_X, FIELD_VOL_FLUX_Y, FIELD_MASS_FLUX_X, FIELD_MASS_FLUX_Y, NUM_FIELDS = 15
};

typedef double* dp;
typedef int* ip;

// ---- the 22 kernel symbols + optional extension (include/clover_b200.h) ----
struct Backend {
  void* handle = nullptr;
  void (*ideal_gas)(ip, ip, ip, ip, dp, dp, dp, dp) = nullptr;
  void (*viscosity)(ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp) = nullptr;
  void (*calc_dt)(ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp,
                  dp, dp, dp, dp, dp, dp, dp, ip, dp, dp, ip, ip, ip) = nullptr;
  void (*pdv)(ip, ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp) = nullptr;
  void (*revert)(ip, ip, ip, ip, dp, dp, dp, dp) = nullptr;
  void (*accelerate)(ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp) = nullptr;
  void (*flux_calc)(ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp, dp) = nullptr;
  void (*advec_cell)(ip, ip, ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp,
                     dp, dp) = nullptr;
  void (*advec_mom)(ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, ip,
                    ip, ip) = nullptr;
  void (*reset_field)(ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp) = nullptr;
  void (*update_halo)(ip, ip, ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp,
                      dp, dp, ip, ip) = nullptr;
  void (*field_summary)(ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp) = nullptr;
  void (*initialise_chunk)(ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp,
                           dp) = nullptr;
  void (*generate_chunk)(ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp, ip, dp, dp, dp, dp, dp, dp,
                         dp, dp, dp, ip, ip, ip, ip) = nullptr;
  // pack/unpack: [face][0=pack,1=unpack]
  void (*packer[4][2])(ip, ip, ip, ip, dp, dp, ip, ip, ip, ip, ip, ip, ip) = {};
  // optional extension
  void (*x_register_chunk)(ip, ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp,
                           dp, dp) = nullptr;
  void (*x_exchange)(ip, ip) = nullptr;
  void (*x_min)(dp) = nullptr;
  void (*x_sum)(dp, ip) = nullptr;
  void (*x_sync_to_host)(ip) = nullptr;
  void (*x_download)(dp) = nullptr;  // one array device -> host (GPU backends only)
  void (*x_forget)(dp) = nullptr;
};

template <class F>
bool sym(void* h, const char* name, F& f, bool required, std::string& err) {
  void* p = dlsym(h, name);
  if (!p) {
    if (required) err += std::string(" missing symbol ") + name + ";";
    return false;
  }
  f = reinterpret_cast<F>(p);
  return true;
}

struct State {
  bool defined = false;
  double density = 0, energy = 0, xvel = 0, yvel = 0;
  double xmin = 0, xmax = 0, ymin = 0, ymax = 0, radius = 0;
  int geometry = 0;
};

struct Deck {  // read_input.f90:38-71 defaults
  double xmin = 0.0, ymin = 0.0, xmax = 100.0, ymax = 100.0;
  int x_cells = 10, y_cells = 10;
  double end_time = 10.0;
  int end_step = g_ibig;
  int visit_frequency = 0, summary_frequency = 10, tiles_per_chunk = 1;
  double dtinit = 0.1, dtmax = 1.0, dtmin = 0.0000001, dtrise = 1.5;
  double dtc_safe = 0.7, dtu_safe = 0.5, dtv_safe = 0.5, dtdiv_safe = 0.7;
  int test_problem = 0;
  bool profiler_on = false;
  std::vector<State> states;  // 1-based: states[0] unused
};

struct Chunk {
  int id = 0;  // 1-based chunk number
  int left = 0, right = 0, bottom = 0, top = 0;
  int x_min = 1, x_max = 0, y_min = 1, y_max = 0;
  int neighbours[4] = {-1, -1, -1, -1};
  int tile_neighbours[4] = {-1, -1, -1, -1};  // one tile per chunk: all external
  // field_type, definitions.f90:125-156
  dp density0 = 0, density1 = 0, energy0 = 0, energy1 = 0, pressure = 0, viscosity = 0, soundspeed = 0;
  dp xvel0 = 0, xvel1 = 0, yvel0 = 0, yvel1 = 0;
  dp vol_flux_x = 0, mass_flux_x = 0, vol_flux_y = 0, mass_flux_y = 0;
  dp work_array1 = 0, work_array2 = 0, work_array3 = 0, work_array4 = 0, work_array5 = 0,
     work_array6 = 0, work_array7 = 0;
  dp cellx = 0, celly = 0, vertexx = 0, vertexy = 0, celldx = 0, celldy = 0, vertexdx = 0, vertexdy = 0;
  dp volume = 0, xarea = 0, yarea = 0;
  dp snd[4] = {}, rcv[4] = {};
  std::vector<void*> allocs;
  dp alloc(size_t n) {
    void* p = calloc(n ? n : 1, sizeof(double));  // build_field.f90:96-181 zero-fills
    if (!p) { fprintf(stderr, "clover_driver: out of host memory\n"); abort(); }
    allocs.push_back(p);
    return (dp)p;
  }
  void release() {
    for (void* p : allocs) free(p);
    allocs.clear();
  }
};

struct StepRec { int step; double time_before, dt; };
struct SummaryRec { int step; double time, vol, mass, density, pressure, ie, ke, total; };

std::string lower_clean(const std::string& in) {
  std::string l = in.substr(0, 100);  // FMT='(a100)'
  for (char& c : l) {
    unsigned char u = (unsigned char)c;
    if (u < 32 || u > 128) c = ' ';
  }
  size_t s = l.find('!');
  if (s != std::string::npos) l = l.substr(0, s);
  s = l.find(';');
  if (s != std::string::npos) l = l.substr(0, s);
  for (char& c : l) {
    if (c >= 'A' && c <= 'Z') c = (char)(c + 32);
    if (c == '=' || c == ',') c = ' ';
  }
  return l;
}

std::vector<std::string> words_of(const std::string& l) {
  std::vector<std::string> w;
  size_t i = 0;
  while (i < l.size()) {
    while (i < l.size() && l[i] == ' ') ++i;
    size_t b = i;
    while (i < l.size() && l[i] != ' ') ++i;
    if (i > b) w.push_back(l.substr(b, i - b));
  }
  return w;
}

}  // namespace

struct clover_driver {
  Deck deck;
  Backend be;
  std::string error;
  int nchunks = 1, rank = 0, comm_mode = 0, verbose = 0;
  int chunk_x = 1, chunk_y = 1;
  std::vector<Chunk> chunks;  // chunks owned by this process
  double time = 0, dt = 0, dtold = 0;
  int step = 0;
  bool advect_x = true, complete = false;
  std::vector<StepRec> steps;
  std::vector<SummaryRec> summaries;
  FILE* out = nullptr;
  double wall_hydro = 0;
  // comm_mode 2: message passing supplied by the host program (MPI in the Fortran original)
  void (*cb_sendrecv)(int peer, const double* snd, double* rcv, int count) = nullptr;
  void (*cb_allreduce)(double* values, int n, int op) = nullptr;  // op 0 = min, 1 = sum

  void log(const char* fmt, ...) {
    if (!out) return;
    va_list ap;
    va_start(ap, fmt);
    vfprintf(out, fmt, ap);
    va_end(ap);
  }

  bool load_backend(const char* path);
  bool parse_deck(const std::string& text);
  void decompose();
  void build_chunk(Chunk& c);
  void start();
  void ideal_gas(Chunk& c, bool predict);
  void update_halo(const int* fields, int depth);
  void exchange(const int* fields, int depth);
  void timestep();
  void pdv(bool predict);
  void accelerate();
  void flux_calc();
  void advec_cell(int sweep, int dir);
  void advec_mom(int which_vel, int dir, int sweep);
  void advection();
  void reset_field();
  void field_summary();
  void visit();
  bool hydro_step();
  std::string visit_dir;  // where visit() writes; empty = visit output disabled (the deck's visit_frequency still parses)
  bool visit_first_call = true;
};

bool clover_driver::load_backend(const char* path) {
  be.handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!be.handle) {
    error = std::string("dlopen failed: ") + dlerror();
    return false;
  }
  void* h = be.handle;
  std::string e;
  sym(h, "ideal_gas_kernel_c_", be.ideal_gas, true, e);
  sym(h, "viscosity_kernel_c_", be.viscosity, true, e);
  sym(h, "calc_dt_kernel_c_", be.calc_dt, true, e);
  sym(h, "pdv_kernel_c_", be.pdv, true, e);
  sym(h, "revert_kernel_c_", be.revert, true, e);
  sym(h, "accelerate_kernel_c_", be.accelerate, true, e);
  sym(h, "flux_calc_kernel_c_", be.flux_calc, true, e);
  sym(h, "advec_cell_kernel_c_", be.advec_cell, true, e);
  sym(h, "advec_mom_kernel_c_", be.advec_mom, true, e);
  sym(h, "reset_field_kernel_c_", be.reset_field, true, e);
  sym(h, "update_halo_kernel_c_", be.update_halo, true, e);
  sym(h, "field_summary_kernel_c_", be.field_summary, true, e);
  sym(h, "initialise_chunk_kernel_c_", be.initialise_chunk, true, e);
  sym(h, "generate_chunk_kernel_c_", be.generate_chunk, true, e);
  static const char* face[4] = {"left", "right", "bottom", "top"};
  for (int f = 0; f < 4; ++f) {
    std::string p = std::string("clover_pack_message_") + face[f] + "_c_";
    std::string u = std::string("clover_unpack_message_") + face[f] + "_c_";
    sym(h, p.c_str(), be.packer[f][0], true, e);
    sym(h, u.c_str(), be.packer[f][1], true, e);
  }
  std::string ignore;
  sym(h, "clover_b200_register_chunk_", be.x_register_chunk, false, ignore);
  sym(h, "clover_b200_exchange_", be.x_exchange, false, ignore);
  sym(h, "clover_b200_min_", be.x_min, false, ignore);
  sym(h, "clover_b200_sum_", be.x_sum, false, ignore);
  sym(h, "clover_b200_sync_to_host_", be.x_sync_to_host, false, ignore);
  sym(h, "clover_b200_download_", be.x_download, false, ignore);
  sym(h, "clover_b200_forget_", be.x_forget, false, ignore);
  if (!e.empty()) {
    error = "backend " + std::string(path) + ":" + e;
    return false;
  }
  if (comm_mode == 1 && !(be.x_register_chunk && be.x_exchange && be.x_min && be.x_sum)) {
    error = "comm_mode=1 needs the clover_b200_* extension symbols";
    return false;
  }
  return true;
}

bool clover_driver::parse_deck(const std::string& text) {
  // Pass 0: collect the lines between *clover and *endclover, cleaned.
  std::vector<std::vector<std::string>> lines;
  bool inside = false;
  size_t pos = 0;
  while (pos <= text.size()) {
    size_t nl = text.find('\n', pos);
    if (nl == std::string::npos) nl = text.size();
    std::string l = lower_clean(text.substr(pos, nl - pos));
    pos = nl + 1;
    std::vector<std::string> w = words_of(l);
    if (w.empty()) continue;
    if (w[0] == "*clover") { inside = true; continue; }
    if (w[0] == "*endclover") { inside = false; continue; }
    if (inside) lines.push_back(w);
  }
  int state_max = 0;
  for (auto& w : lines)
    for (size_t i = 0; i + 1 < w.size(); ++i)
      if (w[i] == "state") { state_max = std::max(state_max, atoi(w[i + 1].c_str())); break; }
  if (state_max < 1) { error = "read_input: No states defined."; return false; }
  deck.states.assign(state_max + 1, State());
  for (auto& w : lines) {
    size_t i = 0;
    auto next = [&]() -> std::string { return (i < w.size()) ? w[i++] : std::string(); };
    while (i < w.size()) {
      std::string word = next();
      if (word == "initial_timestep") deck.dtinit = atof(next().c_str());
      else if (word == "max_timestep") deck.dtmax = atof(next().c_str());
      else if (word == "timestep_rise") deck.dtrise = atof(next().c_str());
      else if (word == "end_time") deck.end_time = atof(next().c_str());
      else if (word == "end_step") deck.end_step = atoi(next().c_str());
      else if (word == "xmin") deck.xmin = atof(next().c_str());
      else if (word == "xmax") deck.xmax = atof(next().c_str());
      else if (word == "ymin") deck.ymin = atof(next().c_str());
      else if (word == "ymax") deck.ymax = atof(next().c_str());
      else if (word == "x_cells") deck.x_cells = atoi(next().c_str());
      else if (word == "y_cells") deck.y_cells = atoi(next().c_str());
      else if (word == "visit_frequency") deck.visit_frequency = atoi(next().c_str());
      else if (word == "summary_frequency") deck.summary_frequency = atoi(next().c_str());
      else if (word == "tiles_per_chunk") deck.tiles_per_chunk = atoi(next().c_str());
      else if (word == "tiles_per_problem") deck.tiles_per_chunk = atoi(next().c_str()) / nchunks;
      else if (word == "profiler_on") deck.profiler_on = true;
      else if (word == "test_problem") deck.test_problem = atoi(next().c_str());
      else if (word == "use_fortran_kernels" || word == "use_c_kernels" || word == "use_oa_kernels") {
        // backend is chosen by the shared library handed to clover_driver_create
      } else if (word == "state") {
        int s = atoi(next().c_str());
        if (s < 1 || s > state_max) { error = "read_input: bad state number"; return false; }
        State& st = deck.states[s];
        if (st.defined) { error = "read_input: State defined twice."; return false; }
        st.defined = true;
        while (i < w.size()) {
          std::string k = next();
          if (k == "xvel") st.xvel = atof(next().c_str());
          else if (k == "yvel") st.yvel = atof(next().c_str());
          else if (k == "xmin") st.xmin = atof(next().c_str());
          else if (k == "ymin") st.ymin = atof(next().c_str());
          else if (k == "xmax") st.xmax = atof(next().c_str());
          else if (k == "ymax") st.ymax = atof(next().c_str());
          else if (k == "radius") st.radius = atof(next().c_str());
          else if (k == "density") st.density = atof(next().c_str());
          else if (k == "energy") st.energy = atof(next().c_str());
          else if (k == "geometry") {
            std::string g = next();
            if (g == "rectangle") st.geometry = G_RECT;
            else if (g == "circle") st.geometry = G_CIRC;
            else if (g == "point") st.geometry = G_POINT;
          }
        }
      }
    }
  }
  if (deck.tiles_per_chunk != 1) {
    error = "tiles_per_chunk != 1 is out of scope (one tile per chunk per GPU)";
    return false;
  }
  // read_input.f90:274-281
  double dx = (deck.xmax - deck.xmin) / (double)(float)deck.x_cells;
  double dy = (deck.ymax - deck.ymin) / (double)(float)deck.y_cells;
  for (int n = 2; n <= state_max; ++n) {
    deck.states[n].xmin += dx / 100.0;
    deck.states[n].ymin += dy / 100.0;
    deck.states[n].xmax -= dx / 100.0;
    deck.states[n].ymax -= dy / 100.0;
  }
  return true;
}

void clover_driver::decompose() {
  // clover.f90:127-196
  const int n = nchunks;
  const int x_cells = deck.x_cells, y_cells = deck.y_cells;
  float mesh_ratio = (float)x_cells / (float)y_cells;  // real()/real(): default REAL
  chunk_x = n;
  chunk_y = 1;
  int split_found = 0;
  for (int c = 1; c <= n; ++c) {
    if (n % c == 0) {
      double factor_x = (double)((float)n / (float)c);
      double factor_y = (double)c;
      if (factor_x / factor_y <= (double)mesh_ratio) {
        chunk_y = c;
        chunk_x = n / c;
        split_found = 1;
        break;
      }
    }
  }
  if (split_found == 0 || chunk_y == n) {
    if (mesh_ratio >= 1.0f) { chunk_x = n; chunk_y = 1; }
    else { chunk_x = 1; chunk_y = n; }
  }
  int delta_x = x_cells / chunk_x, delta_y = y_cells / chunk_y;
  int mod_x = x_cells % chunk_x, mod_y = y_cells % chunk_y;
  int add_x_prev = 0, add_y_prev = 0, cnk = 1;
  chunks.clear();
  for (int cy = 1; cy <= chunk_y; ++cy) {
    for (int cx = 1; cx <= chunk_x; ++cx) {
      int add_x = (cx <= mod_x) ? 1 : 0, add_y = (cy <= mod_y) ? 1 : 0;
      bool mine = (comm_mode == 0) || (cnk == rank + 1);  // modes 1,2: one chunk per process (rank = chunk-1)
      if (mine) {
        Chunk c;
        c.id = cnk;
        c.left = (cx - 1) * delta_x + 1 + add_x_prev;
        c.right = c.left + delta_x - 1 + add_x;
        c.bottom = (cy - 1) * delta_y + 1 + add_y_prev;
        c.top = c.bottom + delta_y - 1 + add_y;
        c.neighbours[LEFT] = chunk_x * (cy - 1) + cx - 1;
        c.neighbours[RIGHT] = chunk_x * (cy - 1) + cx + 1;
        c.neighbours[BOTTOM] = chunk_x * (cy - 2) + cx;
        c.neighbours[TOP] = chunk_x * cy + cx;
        if (cx == 1) c.neighbours[LEFT] = -1;
        if (cx == chunk_x) c.neighbours[RIGHT] = -1;
        if (cy == 1) c.neighbours[BOTTOM] = -1;
        if (cy == chunk_y) c.neighbours[TOP] = -1;
        c.x_min = 1;
        c.y_min = 1;
        c.x_max = c.right - c.left + 1;
        c.y_max = c.top - c.bottom + 1;
        chunks.push_back(c);
      }
      if (cx <= mod_x) add_x_prev++;
      cnk++;
    }
    add_x_prev = 0;
    if (cy <= mod_y) add_y_prev++;
  }
  log("\n Mesh ratio of %g\n Decomposing the mesh into %d by %d chunks\n\n", (double)mesh_ratio,
      chunk_x, chunk_y);
}

void clover_driver::build_chunk(Chunk& c) {
  // build_field.f90:33-94 shapes; clover.f90:329-342 buffers
  const size_t nx = c.x_max, ny = c.y_max;
  const size_t cell = (nx + 4) * (ny + 4), vert = (nx + 5) * (ny + 5);
  const size_t xf = (nx + 5) * (ny + 4), yf = (nx + 4) * (ny + 5);
  c.density0 = c.alloc(cell); c.density1 = c.alloc(cell);
  c.energy0 = c.alloc(cell); c.energy1 = c.alloc(cell);
  c.pressure = c.alloc(cell); c.viscosity = c.alloc(cell); c.soundspeed = c.alloc(cell);
  c.xvel0 = c.alloc(vert); c.xvel1 = c.alloc(vert); c.yvel0 = c.alloc(vert); c.yvel1 = c.alloc(vert);
  c.vol_flux_x = c.alloc(xf); c.mass_flux_x = c.alloc(xf);
  c.vol_flux_y = c.alloc(yf); c.mass_flux_y = c.alloc(yf);
  c.work_array1 = c.alloc(vert); c.work_array2 = c.alloc(vert); c.work_array3 = c.alloc(vert);
  c.work_array4 = c.alloc(vert); c.work_array5 = c.alloc(vert); c.work_array6 = c.alloc(vert);
  c.work_array7 = c.alloc(vert);
  c.cellx = c.alloc(nx + 4); c.celly = c.alloc(ny + 4);
  c.vertexx = c.alloc(nx + 5); c.vertexy = c.alloc(ny + 5);
  c.celldx = c.alloc(nx + 4); c.celldy = c.alloc(ny + 4);
  c.vertexdx = c.alloc(nx + 5); c.vertexdy = c.alloc(ny + 5);
  c.volume = c.alloc(cell); c.xarea = c.alloc(xf); c.yarea = c.alloc(yf);
  for (int f = 0; f < 4; ++f) {
    size_t n = 10 * 2 * ((f < 2 ? ny : nx) + 5);
    c.snd[f] = c.alloc(n);
    c.rcv[f] = c.alloc(n);
  }
}

void clover_driver::ideal_gas(Chunk& c, bool predict) {
  // ideal_gas.f90:62-83
  if (!predict)
    be.ideal_gas(&c.x_min, &c.x_max, &c.y_min, &c.y_max, c.density0, c.energy0, c.pressure, c.soundspeed);
  else
    be.ideal_gas(&c.x_min, &c.x_max, &c.y_min, &c.y_max, c.density1, c.energy1, c.pressure, c.soundspeed);
}

void clover_driver::start() {
  // start.f90:48-143
  log(" Setting up initial geometry\n\n");
  time = 0.0;
  step = 0;
  dtold = deck.dtinit;
  dt = deck.dtinit;
  decompose();
  const double dx = (deck.xmax - deck.xmin) / (double)(float)deck.x_cells;
  const double dy = (deck.ymax - deck.ymin) / (double)(float)deck.y_cells;
  int nstates = (int)deck.states.size() - 1;
  std::vector<double> s_density(nstates), s_energy(nstates), s_xvel(nstates), s_yvel(nstates),
      s_xmin(nstates), s_xmax(nstates), s_ymin(nstates), s_ymax(nstates), s_radius(nstates);
  std::vector<int> s_geom(nstates);
  for (int s = 0; s < nstates; ++s) {
    const State& st = deck.states[s + 1];
    s_density[s] = st.density; s_energy[s] = st.energy; s_xvel[s] = st.xvel; s_yvel[s] = st.yvel;
    s_xmin[s] = st.xmin; s_xmax[s] = st.xmax; s_ymin[s] = st.ymin; s_ymax[s] = st.ymax;
    s_radius[s] = st.radius; s_geom[s] = st.geometry;
  }
  int g_rect = G_RECT, g_circ = G_CIRC, g_point = G_POINT;
  log(" Generating chunks\n");
  for (Chunk& c : chunks) {
    build_chunk(c);
    if (comm_mode == 1)
      be.x_register_chunk(&c.x_min, &c.x_max, &c.y_min, &c.y_max, c.neighbours, c.density0,
                          c.density1, c.energy0, c.energy1, c.pressure, c.viscosity, c.soundspeed,
                          c.xvel0, c.xvel1, c.yvel0, c.yvel1, c.vol_flux_x, c.vol_flux_y,
                          c.mass_flux_x, c.mass_flux_y);
    // initialise_chunk.f90:33-38 (t_left==chunk%left with one tile)
    double xmin = deck.xmin + dx * (double)(float)(c.left - 1);
    double ymin = deck.ymin + dy * (double)(float)(c.bottom - 1);
    double ddx = dx, ddy = dy;
    be.initialise_chunk(&c.x_min, &c.x_max, &c.y_min, &c.y_max, &xmin, &ymin, &ddx, &ddy, c.vertexx,
                        c.vertexdx, c.vertexy, c.vertexdy, c.cellx, c.celldx, c.celly, c.celldy,
                        c.volume, c.xarea, c.yarea);
    be.generate_chunk(&c.x_min, &c.x_max, &c.y_min, &c.y_max, c.vertexx, c.vertexy, c.cellx, c.celly,
                      c.density0, c.energy0, c.xvel0, c.yvel0, &nstates, s_density.data(),
                      s_energy.data(), s_xvel.data(), s_yvel.data(), s_xmin.data(), s_xmax.data(),
                      s_ymin.data(), s_ymax.data(), s_radius.data(), s_geom.data(), &g_rect, &g_circ,
                      &g_point);
  }
  advect_x = true;
  for (Chunk& c : chunks) ideal_gas(c, false);
  int fields[NUM_FIELDS] = {0};
  fields[FIELD_DENSITY0 - 1] = 1; fields[FIELD_ENERGY0 - 1] = 1; fields[FIELD_PRESSURE - 1] = 1;
  fields[FIELD_VISCOSITY - 1] = 1; fields[FIELD_DENSITY1 - 1] = 1; fields[FIELD_ENERGY1 - 1] = 1;
  fields[FIELD_XVEL0 - 1] = 1; fields[FIELD_YVEL0 - 1] = 1; fields[FIELD_XVEL1 - 1] = 1;
  fields[FIELD_YVEL1 - 1] = 1;
  update_halo(fields, 2);
  log("\n Problem initialised and generated\n");
  field_summary();
  if (deck.visit_frequency != 0) visit();  // start.f90:145
}

void clover_driver::exchange(const int* fields, int depth) {
  // clover.f90:348-500
  if (nchunks == 1) return;
  if (comm_mode == 1) {
    int d = depth;
    be.x_exchange(const_cast<int*>(fields), &d);
    return;
  }
  static const int field_type[NUM_FIELDS] = {  // clover.f90:690-880
      CELL_DATA, CELL_DATA, CELL_DATA, CELL_DATA, CELL_DATA, CELL_DATA, CELL_DATA,
      VERTEX_DATA, VERTEX_DATA, VERTEX_DATA, VERTEX_DATA,
      X_FACE_DATA, Y_FACE_DATA, X_FACE_DATA, Y_FACE_DATA};
  int cd = CELL_DATA, vd = VERTEX_DATA, xd = X_FACE_DATA, yd = Y_FACE_DATA, d = depth;
  auto field_ptr = [](Chunk& c, int f) -> dp {
    switch (f + 1) {
      case FIELD_DENSITY0: return c.density0;
      case FIELD_DENSITY1: return c.density1;
      case FIELD_ENERGY0: return c.energy0;
      case FIELD_ENERGY1: return c.energy1;
      case FIELD_PRESSURE: return c.pressure;
      case FIELD_VISCOSITY: return c.viscosity;
      case FIELD_SOUNDSPEED: return c.soundspeed;
      case FIELD_XVEL0: return c.xvel0;
      case FIELD_XVEL1: return c.xvel1;
      case FIELD_YVEL0: return c.yvel0;
      case FIELD_YVEL1: return c.yvel1;
      case FIELD_VOL_FLUX_X: return c.vol_flux_x;
      case FIELD_VOL_FLUX_Y: return c.vol_flux_y;
      case FIELD_MASS_FLUX_X: return c.mass_flux_x;
      case FIELD_MASS_FLUX_Y: return c.mass_flux_y;
    }
    return nullptr;
  };
  if (comm_mode == 2) {
    // The reference's own scheme (clover.f90:348-500): pack into HOST buffers with the backend's
    // pack kernels, exchange them by message passing (callback = MPI_ISEND/IRECV+WAITALL), unpack.
    Chunk& c = chunks[0];
    for (int phase = 0; phase < 2; ++phase) {
      const int fa = phase == 0 ? LEFT : BOTTOM, fb = phase == 0 ? RIGHT : TOP;
      const int edge = (phase == 0 ? c.y_max : c.x_max) + 5;
      int nf = 0;
      for (int f = 0; f < NUM_FIELDS; ++f) nf += (fields[f] == 1);
      for (int face = fa; face <= fb; ++face) {
        if (c.neighbours[face] == -1) continue;
        int off = 0;
        for (int f = 0; f < NUM_FIELDS; ++f) {
          if (fields[f] != 1) continue;
          int ft = field_type[f], o = off;
          be.packer[face][0](&c.x_min, &c.x_max, &c.y_min, &c.y_max, field_ptr(c, f), c.snd[face], &cd, &vd, &xd,
                             &yd, &d, &ft, &o);
          off += depth * edge;
        }
      }
      for (int face = fa; face <= fb; ++face)
        if (c.neighbours[face] != -1)
          cb_sendrecv(c.neighbours[face] - 1, c.snd[face], c.rcv[face], nf * depth * edge);
      for (int face = fa; face <= fb; ++face) {
        if (c.neighbours[face] == -1) continue;
        int off = 0;
        for (int f = 0; f < NUM_FIELDS; ++f) {
          if (fields[f] != 1) continue;
          int ft = field_type[f], o = off;
          be.packer[face][1](&c.x_min, &c.x_max, &c.y_min, &c.y_max, field_ptr(c, f), c.rcv[face], &cd, &vd, &xd,
                             &yd, &d, &ft, &o);
          off += depth * edge;
        }
      }
    }
    return;
  }
  auto by_id = [&](int id) -> Chunk& {
    for (Chunk& c : chunks)
      if (c.id == id) return c;
    fprintf(stderr, "clover_driver: neighbour chunk %d not local\n", id);
    abort();
  };
  for (int phase = 0; phase < 2; ++phase) {
    const int fa = phase == 0 ? LEFT : BOTTOM, fb = phase == 0 ? RIGHT : TOP;
    // pack (per-field offsets: running sum of depth*(edge+5), clover.f90:368-375)
    for (Chunk& c : chunks) {
      int edge = (phase == 0 ? c.y_max : c.x_max) + 5;
      for (int face = fa; face <= fb; ++face) {
        if (c.neighbours[face] == -1) continue;
        int off = 0;
        for (int f = 0; f < NUM_FIELDS; ++f) {
          if (fields[f] != 1) continue;
          int ft = field_type[f], o = off;
          be.packer[face][0](&c.x_min, &c.x_max, &c.y_min, &c.y_max, field_ptr(c, f), c.snd[face],
                             &cd, &vd, &xd, &yd, &d, &ft, &o);
          off += depth * edge;
        }
      }
    }
    // "MPI": my rcv[face] <- neighbour's snd[opposite face]
    for (Chunk& c : chunks) {
      int edge = (phase == 0 ? c.y_max : c.x_max) + 5;
      int nf = 0;
      for (int f = 0; f < NUM_FIELDS; ++f) nf += (fields[f] == 1);
      size_t total = (size_t)nf * depth * edge;
      for (int face = fa; face <= fb; ++face) {
        if (c.neighbours[face] == -1) continue;
        Chunk& nb = by_id(c.neighbours[face]);
        int opp = (face == LEFT) ? RIGHT : (face == RIGHT) ? LEFT : (face == BOTTOM) ? TOP : BOTTOM;
        memcpy(c.rcv[face], nb.snd[opp], total * sizeof(double));
      }
    }
    // unpack
    for (Chunk& c : chunks) {
      int edge = (phase == 0 ? c.y_max : c.x_max) + 5;
      for (int face = fa; face <= fb; ++face) {
        if (c.neighbours[face] == -1) continue;
        int off = 0;
        for (int f = 0; f < NUM_FIELDS; ++f) {
          if (fields[f] != 1) continue;
          int ft = field_type[f], o = off;
          be.packer[face][1](&c.x_min, &c.x_max, &c.y_min, &c.y_max, field_ptr(c, f), c.rcv[face],
                             &cd, &vd, &xd, &yd, &d, &ft, &o);
          off += depth * edge;
        }
      }
    }
  }
}

void clover_driver::update_halo(const int* fields, int depth) {
  // update_halo.f90:39-113 (update_tile_halo is a no-op with one tile per chunk)
  exchange(fields, depth);
  int d = depth;
  for (Chunk& c : chunks) {
    if (c.neighbours[LEFT] == -1 || c.neighbours[RIGHT] == -1 || c.neighbours[BOTTOM] == -1 ||
        c.neighbours[TOP] == -1) {
      be.update_halo(&c.x_min, &c.x_max, &c.y_min, &c.y_max, c.neighbours, c.tile_neighbours,
                     c.density0, c.energy0, c.pressure, c.viscosity, c.soundspeed, c.density1,
                     c.energy1, c.xvel0, c.yvel0, c.xvel1, c.yvel1, c.vol_flux_x, c.vol_flux_y,
                     c.mass_flux_x, c.mass_flux_y, const_cast<int*>(fields), &d);
    }
  }
}

void clover_driver::timestep() {
  // timestep.f90:56-117
  dt = g_big;
  for (Chunk& c : chunks) ideal_gas(c, false);
  int fields[NUM_FIELDS] = {0};
  fields[FIELD_PRESSURE - 1] = 1; fields[FIELD_ENERGY0 - 1] = 1; fields[FIELD_DENSITY0 - 1] = 1;
  fields[FIELD_XVEL0 - 1] = 1; fields[FIELD_YVEL0 - 1] = 1;
  update_halo(fields, 1);
  for (Chunk& c : chunks)  // viscosity.f90:57
    be.viscosity(&c.x_min, &c.x_max, &c.y_min, &c.y_max, c.celldx, c.celldy, c.density0, c.pressure,
                 c.viscosity, c.xvel0, c.yvel0);
  memset(fields, 0, sizeof(fields));
  fields[FIELD_VISCOSITY - 1] = 1;
  update_halo(fields, 1);
  int dt_control = 1, jdt = 0, kdt = 0;
  double x_pos = 0, y_pos = 0;
  for (Chunk& c : chunks) {  // calc_dt.f90:83
    double gs = g_small, gb = g_big, dtmin = deck.dtmin, dtc = deck.dtc_safe, dtu = deck.dtu_safe,
           dtv = deck.dtv_safe, dtdiv = deck.dtdiv_safe;
    double dtlp = g_big, xl = 0, yl = 0;
    int ctl = 0, jl = 0, kl = 0, small = 0;
    be.calc_dt(&c.x_min, &c.x_max, &c.y_min, &c.y_max, &gs, &gb, &dtmin, &dtc, &dtu, &dtv, &dtdiv,
               c.xarea, c.yarea, c.cellx, c.celly, c.celldx, c.celldy, c.volume, c.density0,
               c.energy0, c.pressure, c.viscosity, c.soundspeed, c.xvel0, c.yvel0, c.work_array1,
               &dtlp, &ctl, &xl, &yl, &jl, &kl, &small);
    if (dtlp <= dt) { dt = dtlp; dt_control = ctl; x_pos = xl; y_pos = yl; jdt = jl; kdt = kl; }
  }
  dt = std::min(dt, std::min(dtold * deck.dtrise, deck.dtmax));
  if (comm_mode == 1 && nchunks > 1) be.x_min(&dt);  // clover_min, clover.f90:3653
  if (comm_mode == 2 && nchunks > 1) cb_allreduce(&dt, 1, 0);
  static const char* names[5] = {"", "sound", "xvel", "yvel", "div"};
  log(" Step %7d time %11.7f control %11s timestep  %9.2E%8d,%8d x %9.2E y %9.2E\n", step, time,
      names[(dt_control >= 1 && dt_control <= 4) ? dt_control : 0], dt, jdt, kdt, x_pos, y_pos);
  if (dt < deck.dtmin) {
    error = "timestep: small timestep";
    complete = true;
  }
  steps.push_back({step, time, dt});
  dtold = dt;
}

void clover_driver::pdv(bool predict) {
  // PdV.f90:46-138
  int prdct = predict ? 0 : 1;
  for (Chunk& c : chunks)
    be.pdv(&prdct, &c.x_min, &c.x_max, &c.y_min, &c.y_max, &dt, c.xarea, c.yarea, c.volume,
           c.density0, c.density1, c.energy0, c.energy1, c.pressure, c.viscosity, c.xvel0, c.xvel1,
           c.yvel0, c.yvel1, c.work_array1);
  // clover_check_error(error_condition): error_condition is never set (PdV.f90:46) -- dropped.
  if (predict) {
    for (Chunk& c : chunks) ideal_gas(c, true);
    int fields[NUM_FIELDS] = {0};
    fields[FIELD_PRESSURE - 1] = 1;
    update_halo(fields, 1);
    for (Chunk& c : chunks)  // revert.f90:53
      be.revert(&c.x_min, &c.x_max, &c.y_min, &c.y_max, c.density0, c.density1, c.energy0, c.energy1);
  }
}

void clover_driver::accelerate() {
  for (Chunk& c : chunks)  // accelerate.f90:64
    be.accelerate(&c.x_min, &c.x_max, &c.y_min, &c.y_max, &dt, c.xarea, c.yarea, c.volume, c.density0,
                  c.pressure, c.viscosity, c.xvel0, c.yvel0, c.xvel1, c.yvel1);
}

void clover_driver::flux_calc() {
  for (Chunk& c : chunks)  // flux_calc.f90:62
    be.flux_calc(&c.x_min, &c.x_max, &c.y_min, &c.y_max, &dt, c.xarea, c.yarea, c.xvel0, c.yvel0,
                 c.xvel1, c.yvel1, c.vol_flux_x, c.vol_flux_y);
}

void clover_driver::advec_cell(int sweep, int dir) {
  for (Chunk& c : chunks)  // advec_cell_driver.f90:59
    be.advec_cell(&c.x_min, &c.x_max, &c.y_min, &c.y_max, &dir, &sweep, c.vertexdx, c.vertexdy,
                  c.volume, c.density1, c.energy1, c.mass_flux_x, c.vol_flux_x, c.mass_flux_y,
                  c.vol_flux_y, c.work_array1, c.work_array2, c.work_array3, c.work_array4,
                  c.work_array5, c.work_array6, c.work_array7);
}

void clover_driver::advec_mom(int which_vel, int dir, int sweep) {
  for (Chunk& c : chunks)  // advec_mom_driver.f90:85,108
    be.advec_mom(&c.x_min, &c.x_max, &c.y_min, &c.y_max, which_vel == 1 ? c.xvel1 : c.yvel1,
                 c.mass_flux_x, c.vol_flux_x, c.mass_flux_y, c.vol_flux_y, c.volume, c.density1,
                 c.work_array1, c.work_array2, c.work_array3, c.work_array4, c.work_array5,
                 c.work_array6, c.celldx, c.celldy, &which_vel, &sweep, &dir);
}

void clover_driver::advection() {
  // advection.f90:43-110
  int sweep = 1;
  int dir = advect_x ? 1 : 2;
  int fields[NUM_FIELDS] = {0};
  fields[FIELD_ENERGY1 - 1] = 1; fields[FIELD_DENSITY1 - 1] = 1;
  fields[FIELD_VOL_FLUX_X - 1] = 1; fields[FIELD_VOL_FLUX_Y - 1] = 1;
  update_halo(fields, 2);
  advec_cell(sweep, dir);
  memset(fields, 0, sizeof(fields));
  fields[FIELD_DENSITY1 - 1] = 1; fields[FIELD_ENERGY1 - 1] = 1; fields[FIELD_XVEL1 - 1] = 1;
  fields[FIELD_YVEL1 - 1] = 1; fields[FIELD_MASS_FLUX_X - 1] = 1; fields[FIELD_MASS_FLUX_Y - 1] = 1;
  update_halo(fields, 2);
  advec_mom(1, dir, sweep);
  advec_mom(2, dir, sweep);
  sweep = 2;
  dir = advect_x ? 2 : 1;
  advec_cell(sweep, dir);
  update_halo(fields, 2);
  advec_mom(1, dir, sweep);
  advec_mom(2, dir, sweep);
}

void clover_driver::reset_field() {
  for (Chunk& c : chunks)  // reset_field.f90:63
    be.reset_field(&c.x_min, &c.x_max, &c.y_min, &c.y_max, c.density0, c.density1, c.energy0,
                   c.energy1, c.xvel0, c.xvel1, c.yvel0, c.yvel1);
}

void clover_driver::field_summary() {
  // field_summary.f90:53-129
  for (Chunk& c : chunks) ideal_gas(c, false);
  double t[5] = {0, 0, 0, 0, 0};  // vol, mass, ie, ke, press
  for (Chunk& c : chunks) {
    double vol = 0, mass = 0, ie = 0, ke = 0, press = 0;
    be.field_summary(&c.x_min, &c.x_max, &c.y_min, &c.y_max, c.volume, c.density0, c.energy0,
                     c.pressure, c.xvel0, c.yvel0, &vol, &mass, &ie, &ke, &press);
    t[0] += vol; t[1] += mass; t[2] += ie; t[3] += ke; t[4] += press;
  }
  if (comm_mode == 1 && nchunks > 1) {  // clover_sum x5, clover.f90:3635
    int n = 5;
    be.x_sum(t, &n);
  }
  if (comm_mode == 2 && nchunks > 1) cb_allreduce(t, 5, 1);
  SummaryRec r{step, time, t[0], t[1], t[1] / t[0], t[4] / t[0], t[2], t[3], t[2] + t[3]};
  summaries.push_back(r);
  log("\n Time %.16g\n%13s%16s%16s%16s%16s%16s%16s%16s\n", time, "", "Volume", "Mass", "Density",
      "Pressure", "Internal Energy", "Kinetic Energy", "Total Energy");
  log(" step:%7d%16.4E%16.4E%16.4E%16.4E%16.4E%16.4E%16.4E\n\n", step, r.vol, r.mass, r.density,
      r.pressure, r.ie, r.ke, r.total);
  if (complete && deck.test_problem >= 1 && deck.test_problem <= 5) {
    static const double gold[6] = {0, 1.82280367310258, 1.19316898756307, 2.58984003503994,
                                   0.307475452287895, 4.85350315783719};  // field_summary.f90:139-143
    double qa = std::fabs(100.0 * (r.ke / gold[deck.test_problem]) - 100.0);
    log("Test problem%4d is within%16.7E%% of the expected solution\n", deck.test_problem, qa);
    log(qa < 0.001 ? " This test is considered PASSED\n" : " This test is considered NOT PASSED\n");
  }
}

// Fortran `E12.4` edit descriptor: 0.dddd mantissa, two-digit exponent, right-justified in 12 columns
static void fortran_e12_4(FILE* f, double v) {
  if (v == 0.0 || !std::isfinite(v)) {
    if (v == 0.0) fprintf(f, "%s\n", std::signbit(v) ? " -0.0000E+00" : "  0.0000E+00");
    else fprintf(f, "%12s\n", std::isnan(v) ? "NaN" : (v > 0 ? "Infinity" : "-Infinity"));
    return;
  }
  int e = (int)std::floor(std::log10(std::fabs(v))) + 1;
  double m = std::fabs(v) / std::pow(10.0, e);
  long digits = std::lround(m * 1.0e4);
  if (digits >= 10000) { digits = 1000; e += 1; }   // 0.99996 rounds up to 1.0000 -> 0.1000E+(e+1)
  if (digits < 1000) { digits *= 10; e -= 1; }       // log10 landed one decade high
  char buf[32];
  snprintf(buf, sizeof buf, "%s0.%04ldE%c%02d", v < 0 ? "-" : "", digits, e < 0 ? '-' : '+', e < 0 ? -e : e);
  fprintf(f, "%12s\n", buf);
}

void clover_driver::visit() {
  // visit.f90:25-180: refresh pressure and viscosity, then one ASCII VTK rectilinear-grid file per chunk and
  // an index file `clover.visit`.  Disabled unless an output directory was given (clover_driver_set_visit).
  if (visit_dir.empty()) return;
  const bool boss = (comm_mode == 0) || rank == 0;
  const std::string index = visit_dir + "/clover.visit";
  if (boss && visit_first_call) {  // :52-60
    FILE* u = fopen(index.c_str(), "w");
    if (!u) { error = "visit: cannot write " + index; return; }
    fprintf(u, "!NBLOCKS %5d\n", nchunks * deck.tiles_per_chunk);
    fclose(u);
  }
  visit_first_call = false;
  for (Chunk& c : chunks) ideal_gas(c, false);  // :65-67
  int fields[NUM_FIELDS] = {0};
  fields[FIELD_PRESSURE - 1] = 1; fields[FIELD_XVEL0 - 1] = 1; fields[FIELD_YVEL0 - 1] = 1;
  update_halo(fields, 1);                        // :70-74
  for (Chunk& c : chunks)                         // :77
    be.viscosity(&c.x_min, &c.x_max, &c.y_min, &c.y_max, c.celldx, c.celldy, c.density0, c.pressure,
                 c.viscosity, c.xvel0, c.yvel0);
  // a GPU backend keeps the fields on the device: bring the six dumped ones back (the D2H path of visit)
  if (be.x_download)
    for (Chunk& c : chunks)
      for (dp a : {c.density0, c.energy0, c.pressure, c.viscosity, c.xvel0, c.yvel0, c.vertexx, c.vertexy})
        be.x_download(a);  // vertexx/vertexy: visit.f90:127,131 (already on the host after initialise_chunk; cheap)
  auto vtk_name = [&](int task) {
    char b[64];
    snprintf(b, sizeof b, "clover.%05d.%05d.%05d.vtk", task, 1, step);  // i6 of n+100000 with the '1' -> '.'
    return std::string(b);
  };
  if (boss) {  // :81-99
    FILE* u = fopen(index.c_str(), "a");
    if (!u) { error = "visit: cannot append to " + index; return; }
    for (int task = 0; task < nchunks; ++task) fprintf(u, "%s\n", vtk_name(task).c_str());
    fclose(u);
  }
  for (Chunk& c : chunks) {  // :103-177
    const int nxc = c.x_max - c.x_min + 1, nyc = c.y_max - c.y_min + 1, nxv = nxc + 1, nyv = nyc + 1;
    const std::string fn = visit_dir + "/" + vtk_name(c.id - 1);
    FILE* u = fopen(fn.c_str(), "w");
    if (!u) { error = "visit: cannot write " + fn; return; }
    const size_t rc = (size_t)(c.x_max + 4), rv = (size_t)(c.x_max + 5);  // cell / vertex row lengths (lower bound -1)
    auto cell = [&](dp a, int j, int k) { return a[(size_t)(k + 1) * rc + (size_t)(j + 1)]; };
    auto vert = [&](dp a, int j, int k) { return a[(size_t)(k + 1) * rv + (size_t)(j + 1)]; };
    fprintf(u, "# vtk DataFile Version 3.0\nvtk output\nASCII\nDATASET RECTILINEAR_GRID\n");
    fprintf(u, "DIMENSIONS%12d%12d 1\n", nxv, nyv);
    fprintf(u, "X_COORDINATES %5d double\n", nxv);
    for (int j = c.x_min; j <= c.x_max + 1; ++j) fortran_e12_4(u, c.vertexx[j + 1]);
    fprintf(u, "Y_COORDINATES %5d double\n", nyv);
    for (int k = c.y_min; k <= c.y_max + 1; ++k) fortran_e12_4(u, c.vertexy[k + 1]);
    fprintf(u, "Z_COORDINATES 1 double\n0\n");
    fprintf(u, "CELL_DATA %20d\nFIELD FieldData 4\n", nxc * nyc);
    const struct { const char* name; dp a; bool clip; } cells[4] = {
        {"density", c.density0, false}, {"energy", c.energy0, false}, {"pressure", c.pressure, false},
        {"viscosity", c.viscosity, true}};
    for (const auto& fld : cells) {
      fprintf(u, "%s 1 %20d double\n", fld.name, nxc * nyc);
      for (int k = c.y_min; k <= c.y_max; ++k)
        for (int j = c.x_min; j <= c.x_max; ++j) {
          double v = cell(fld.a, j, k);
          if (fld.clip && !(v > 0.00000001)) v = 0.0;  // :146-147
          fortran_e12_4(u, v);
        }
    }
    fprintf(u, "POINT_DATA %20d\nFIELD FieldData 2\n", nxv * nyv);
    const struct { const char* name; dp a; } verts[2] = {{"x_vel", c.xvel0}, {"y_vel", c.yvel0}};
    for (const auto& fld : verts) {
      fprintf(u, "%s 1 %20d double\n", fld.name, nxv * nyv);
      for (int k = c.y_min; k <= c.y_max + 1; ++k)
        for (int j = c.x_min; j <= c.x_max + 1; ++j) {
          double v = vert(fld.a, j, k);
          if (!(std::fabs(v) > 0.00000001)) v = 0.0;  // :155-157, :164-166
          fortran_e12_4(u, v);
        }
    }
    fclose(u);
  }
}

bool clover_driver::hydro_step() {
  // hydro.f90:48-99 (one trip of the DO loop); returns false once complete
  if (complete) return false;
  step = step + 1;
  timestep();
  if (complete) return false;  // small-timestep abort
  pdv(true);
  accelerate();
  pdv(false);
  flux_calc();
  advection();
  reset_field();
  advect_x = !advect_x;
  time = time + dt;
  if (deck.summary_frequency != 0 && step % deck.summary_frequency == 0) field_summary();
  if (deck.visit_frequency != 0 && step % deck.visit_frequency == 0) visit();  // hydro.f90:73-75
  if (time + g_small > deck.end_time || step >= deck.end_step) {
    complete = true;
    field_summary();
    if (deck.visit_frequency != 0) visit();  // hydro.f90:88
    log("\n Calculation complete\n Clover is finishing\n");
    return false;
  }
  return true;
}

// ------------------------------- C API --------------------------------------
extern "C" {

// comm_mode 0: all `nchunks` chunks in this process (MPI emulated by memcpy);
// comm_mode 1: this process owns chunk `rank`+1, backend does exchange/reductions.
clover_driver* clover_driver_create(const char* deck_text, const char* backend_so, int nchunks,
                                    int rank, int comm_mode, const char* log_path) {
  clover_driver* d = new clover_driver();
  d->nchunks = nchunks < 1 ? 1 : nchunks;
  d->rank = rank;
  d->comm_mode = comm_mode;
  if (log_path && log_path[0]) d->out = fopen(log_path, "w");
  if (!d->parse_deck(deck_text ? deck_text : "")) return d;
  d->load_backend(backend_so);
  return d;
}

const char* clover_driver_error(clover_driver* d) { return d->error.c_str(); }

// comm_mode 2 only: the two message-passing primitives the reference takes from MPI.
void clover_driver_set_comm_callbacks(clover_driver* d, void (*sendrecv)(int, const double*, double*, int),
                                      void (*allreduce)(double*, int, int)) {
  d->cb_sendrecv = sendrecv;
  d->cb_allreduce = allreduce;
}

void clover_driver_set_end_step(clover_driver* d, int end_step) { d->deck.end_step = end_step; }
void clover_driver_set_end_time(clover_driver* d, double end_time) { d->deck.end_time = end_time; }
void clover_driver_set_summary_frequency(clover_driver* d, int f) { d->deck.summary_frequency = f; }
// visit.f90 output: directory for clover.visit / *.vtk and the dump frequency in steps (0 keeps the deck's value)
void clover_driver_set_visit(clover_driver* d, const char* dir, int frequency) {
  d->visit_dir = dir ? dir : "";
  if (frequency > 0) d->deck.visit_frequency = frequency;
}

int clover_driver_start(clover_driver* d) {
  if (!d->error.empty()) return -1;
  if (d->comm_mode == 2 && d->nchunks > 1 && !(d->cb_sendrecv && d->cb_allreduce)) {
    d->error = "comm_mode=2 needs clover_driver_set_comm_callbacks";
    return -1;
  }
  d->start();
  return d->error.empty() ? 0 : -1;
}

// Runs up to n trips of the hydro loop; returns the number actually run.
int clover_driver_run(clover_driver* d, int n) {
  if (!d->error.empty()) return -1;
  int done = 0;
  auto t0 = std::chrono::steady_clock::now();
  while (done < n && !d->complete) {
    d->hydro_step();
    ++done;
  }
  d->wall_hydro += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (d->out) fflush(d->out);
  return done;
}

void clover_driver_field_summary(clover_driver* d) { d->field_summary(); }

int clover_driver_complete(clover_driver* d) { return d->complete ? 1 : 0; }
int clover_driver_step(clover_driver* d) { return d->step; }
double clover_driver_time(clover_driver* d) { return d->time; }
double clover_driver_dt(clover_driver* d) { return d->dt; }
double clover_driver_wall(clover_driver* d) { return d->wall_hydro; }
int clover_driver_num_chunks(clover_driver* d) { return (int)d->chunks.size(); }
void clover_driver_grid(clover_driver* d, int* out4) {
  out4[0] = d->deck.x_cells; out4[1] = d->deck.y_cells; out4[2] = d->chunk_x; out4[3] = d->chunk_y;
}

int clover_driver_num_steps(clover_driver* d) { return (int)d->steps.size(); }
// out: n x 3 doubles (step, time before the step, dt)
void clover_driver_get_steps(clover_driver* d, double* out) {
  for (size_t i = 0; i < d->steps.size(); ++i) {
    out[3 * i] = d->steps[i].step; out[3 * i + 1] = d->steps[i].time_before; out[3 * i + 2] = d->steps[i].dt;
  }
}
int clover_driver_num_summaries(clover_driver* d) { return (int)d->summaries.size(); }
// out: n x 9 doubles (step, time, vol, mass, density, pressure, ie, ke, total)
void clover_driver_get_summaries(clover_driver* d, double* out) {
  for (size_t i = 0; i < d->summaries.size(); ++i) {
    const SummaryRec& r = d->summaries[i];
    double v[9] = {(double)r.step, r.time, r.vol, r.mass, r.density, r.pressure, r.ie, r.ke, r.total};
    memcpy(out + 9 * i, v, sizeof(v));
  }
}

// Chunk geometry: out[0..9] = id,left,right,bottom,top,x_max,y_max? see below
void clover_driver_chunk_info(clover_driver* d, int idx, int* out) {
  const Chunk& c = d->chunks[idx];
  int v[11] = {c.id, c.left, c.right, c.bottom, c.top, c.x_max, c.y_max, c.neighbours[0],
               c.neighbours[1], c.neighbours[2], c.neighbours[3]};
  memcpy(out, v, sizeof(v));
}

// Host pointer of a named field of local chunk idx (for tests; in resident GPU
// mode call clover_b200_sync_to_host_ first).
double* clover_driver_field(clover_driver* d, int idx, const char* name) {
  Chunk& c = d->chunks[idx];
  struct { const char* n; dp p; } tab[] = {
      {"density0", c.density0}, {"density1", c.density1}, {"energy0", c.energy0},
      {"energy1", c.energy1}, {"pressure", c.pressure}, {"viscosity", c.viscosity},
      {"soundspeed", c.soundspeed}, {"xvel0", c.xvel0}, {"xvel1", c.xvel1}, {"yvel0", c.yvel0},
      {"yvel1", c.yvel1}, {"vol_flux_x", c.vol_flux_x}, {"vol_flux_y", c.vol_flux_y},
      {"mass_flux_x", c.mass_flux_x}, {"mass_flux_y", c.mass_flux_y}, {"volume", c.volume},
      {"xarea", c.xarea}, {"yarea", c.yarea}, {"cellx", c.cellx}, {"celly", c.celly},
      {"celldx", c.celldx}, {"celldy", c.celldy}, {"vertexx", c.vertexx}, {"vertexy", c.vertexy},
      {"vertexdx", c.vertexdx}, {"vertexdy", c.vertexdy}};
  for (auto& t : tab)
    if (strcmp(t.n, name) == 0) return t.p;
  return nullptr;
}

void clover_driver_sync_to_host(clover_driver* d) {
  // every local chunk's 15 hydro fields and 2-D geometry, by address (the backend's sync_to_host_ knows only the
  // one chunk registered for exchange, which is not enough with several chunks per process)
  if (d->be.x_download) {
    for (Chunk& c : d->chunks)
      for (dp a : {c.density0, c.density1, c.energy0, c.energy1, c.pressure, c.viscosity, c.soundspeed, c.xvel0,
                   c.xvel1, c.yvel0, c.yvel1, c.vol_flux_x, c.vol_flux_y, c.mass_flux_x, c.mass_flux_y, c.volume,
                   c.xarea, c.yarea})
        d->be.x_download(a);
  } else if (d->be.x_sync_to_host) {
    int fields[NUM_FIELDS];
    for (int i = 0; i < NUM_FIELDS; ++i) fields[i] = 1;
    d->be.x_sync_to_host(fields);
  }
}

void clover_driver_destroy(clover_driver* d) {
  if (!d) return;
  // a GPU backend keys its device mirrors by host address: drop them before the addresses are recycled
  if (d->be.x_forget)
    for (Chunk& c : d->chunks)
      for (void* p : c.allocs) d->be.x_forget((dp)p);
  for (Chunk& c : d->chunks) c.release();
  if (d->out) fclose(d->out);
  // the backend handle is intentionally left open (device state lives in it)
  delete d;
}

}  // extern "C"
