// extern "C" entry points declared in include/jaxib_b200.h: argument checking, scratch carving
// and the fused step driver.  Nothing here synchronises; every call only enqueues on `stream`.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <math.h>

#include <algorithm>
#include <initializer_list>
#include <utility>
#include <vector>

#include "common.cuh"

namespace jaxib {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled(int kernel_bit) {
  static int mask = -1;
  if (mask < 0) {
    const char* m = getenv("JAXIB_PDL_MASK");
    mask = getenv("JAXIB_NO_PDL") ? 0 : (m ? atoi(m) : 0x7fffffff);
  }
  return (mask & kernel_bit) != 0;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return JAXIB_ERR_CUDA;
  }
  return JAXIB_OK;
}

int check_grid(const jaxib_grid* g) {
  if (!g) { set_error("grid is null"); return JAXIB_ERR_INVALID; }
  if (g->nx <= 0 || g->ny <= 0 || g->batch < 1) {
    set_error("grid: nx=%d ny=%d batch=%d must be positive", g->nx, g->ny, g->batch);
    return JAXIB_ERR_INVALID;
  }
  if (!(g->dx > 0.0) || !(g->dy > 0.0)) { set_error("grid: dx, dy must be positive"); return JAXIB_ERR_INVALID; }
  if (g->bc != JAXIB_BC_PERIODIC && g->bc != JAXIB_BC_CHANNEL && g->bc != JAXIB_BC_FAR_FIELD) {
    set_error("grid: unknown bc %d", g->bc);
    return JAXIB_ERR_INVALID;
  }
  if (g->convect < 0 || g->convect > JAXIB_CONVECT_NONE) {
    set_error("grid: unknown convect scheme %d", g->convect);
    return JAXIB_ERR_INVALID;
  }
  return JAXIB_OK;
}

static int check_ptrs(const char* who, std::initializer_list<const void*> ptrs) {
  for (const void* p : ptrs)
    if (!p) { set_error("%s: null device pointer", who); return JAXIB_ERR_INVALID; }
  return JAXIB_OK;
}

// ---- immersed-boundary chain on a side stream --------------------------------------------------------------
// The two IB kernels (marker forces, spreading) are pure latency chains that occupy a few SMs for ~20 us.  When
// the host knows where the bodies are (body_table_host), the predictor is split: rows near the bodies first
// (K1a, a short launch), then the IB chain runs on a high-priority side stream WHILE the predictor covers the
// rest of the grid (K1b); the projection waits for both.  One side stream + two events per device, created on
// first use (process-wide, like the slab decomposition's side stream: documented in the header).
struct OverlapCtx {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool tried = false;
};
static OverlapCtx* overlap_ctx() {
  static OverlapCtx ctx[32];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return nullptr;
  OverlapCtx& c = ctx[dev];
  if (!c.tried) {
    c.tried = true;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&c.side, cudaStreamNonBlocking, hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming) != cudaSuccess)
      c.side = nullptr;
    cudaGetLastError();
  }
  return c.side ? &c : nullptr;
}
static bool overlap_enabled() {
  static int on = -1;
  // Off by default.  Measured on B200 (profiles/r2h_overlap_trace.log, ib2048): the results are bit-identical, but
  // next to the bandwidth-saturating predictor the two latency-bound IB kernels take 30 us instead of 22 us, the
  // short first launch costs ~9 us and each cross-stream hop ~3 us: 92.5 us per step against 86.9 us serial.
  if (on < 0) { const char* e = getenv("JAXIB_IB_OVERLAP"); on = (e && atoi(e) == 1) ? 1 : 0; }
  return on == 1;
}

// Row ranges touched by the IB kernels of one step (interpolation windows and spreading boxes of every body,
// same extent as launch_ib_spread_bodies) and their complement.  Returns false when splitting is not worth it.
static bool body_row_bands(const GridF& g, const jaxib_bodies& b, const float* state_host, RowBands* near, RowBands* rest) {
  const float hmin = g.dx < g.dy ? g.dx : g.dy;
  const int he = (int)ceil(b.max_radius / hmin) + g.R + 2;
  int lo[kMaxBands], hi[kMaxBands], n = 0;
  std::vector<std::pair<int, int>> iv;
  for (int k = 0; k < b.n_bodies; ++k) {
    const float xc = state_host[(size_t)k * JAXIB_BODY_STATE_STRIDE + 2];
    const int ic = (int)floorf((xc - g.lower_x) / g.dx - 1.0f);
    int a = ic - he - 2, e = ic + he + 3;
    if (a < 0) a = 0;
    if (e > g.nx) e = g.nx;
    if (a < e) iv.push_back({a, e});
  }
  if (iv.empty()) return false;
  std::sort(iv.begin(), iv.end());
  for (auto& x : iv) {
    if (n > 0 && x.first <= hi[n - 1]) { if (x.second > hi[n - 1]) hi[n - 1] = x.second; continue; }
    if (n == kMaxBands - 1) return false;
    lo[n] = x.first; hi[n] = x.second; ++n;
  }
  int covered = 0;
  for (int k = 0; k < n; ++k) covered += hi[k] - lo[k];
  if (covered * 2 > g.nx) return false;  // bodies everywhere: nothing to hide the IB chain behind
  near->n = n;
  for (int k = 0; k < n; ++k) { near->start[k] = lo[k]; near->rows[k] = hi[k] - lo[k]; }
  int m = 0, at = 0;
  for (int k = 0; k <= n; ++k) {
    const int end = k < n ? lo[k] : g.nx;
    if (end > at) { rest->start[m] = at; rest->rows[m] = end - at; ++m; }
    if (k < n) at = hi[k];
  }
  rest->n = m;
  return m > 0;
}

struct Scratch {
  float* u_star;
  float* v_star;
  float* spectrum;
  float* markers;
  unsigned* sync;  // kSyncBytes of device words for kernel-to-kernel completion counters (RowDeps)
};
constexpr size_t kSyncBytes = 256;

static size_t scratch_need(const jaxib_plan* plan, int n_markers) {
  const GridF& g = plan_grid(plan);
  size_t plane = (size_t)g.batch * g.nx * g.ny;
  size_t spectrum = spectrum_floats(plan);  // [nx][ny/2 + 4] complex, or the x-sweep's row images + carried values
  return (2 * plane + spectrum + (size_t)5 * g.batch * (n_markers > 0 ? n_markers : 0)) * sizeof(float) + 256 + kSyncBytes;
}

static int carve(const jaxib_plan* plan, int n_markers, void* scratch, size_t bytes, Scratch* out) {
  const GridF& g = plan_grid(plan);
  size_t need = scratch_need(plan, n_markers);
  if (!scratch || bytes < need) {
    set_error("scratch too small: have %zu bytes, need %zu", bytes, need);
    return JAXIB_ERR_SCRATCH;
  }
  uintptr_t a = (reinterpret_cast<uintptr_t>(scratch) + 255) & ~uintptr_t(255);
  float* base = reinterpret_cast<float*>(a);
  size_t plane = (size_t)g.batch * g.nx * g.ny;
  out->u_star = base;
  out->v_star = base + plane;
  out->spectrum = base + 2 * plane;
  out->markers = base + 2 * plane + spectrum_floats(plan);
  out->sync = reinterpret_cast<unsigned*>(
      (reinterpret_cast<uintptr_t>(out->markers + (size_t)5 * g.batch * (n_markers > 0 ? n_markers : 0)) + 63) & ~uintptr_t(63));
  return JAXIB_OK;
}

}  // namespace jaxib

using namespace jaxib;

extern "C" {

const char* jaxib_version(void) { return "jax_ib_b200 0.1 (sm_100a)"; }
const char* jaxib_last_error(void) { return g_err; }
int jaxib_compiled_arch(void) {
#ifdef JAXIB_ARCH
  return JAXIB_ARCH;
#else
  return 100;
#endif
}

size_t jaxib_scratch_bytes(const jaxib_plan* plan, int32_t n_markers) {
  if (!plan) return 0;
  return scratch_need(plan, n_markers);
}

int jaxib_explicit_terms(jaxib_stream_t stream, const jaxib_grid* grid, const float* u, const float* v,
                         const float* perm, float* du, float* dv) {
  int rc = check_grid(grid);
  if (rc) return rc;
  if ((rc = check_ptrs("explicit_terms", {u, v, du, dv}))) return rc;
  return launch_explicit((cudaStream_t)stream, make_gridf(*grid), u, v, nullptr, perm, du, dv, false);
}

int jaxib_explicit_predict(jaxib_stream_t stream, const jaxib_grid* grid, const float* u, const float* v,
                           const float* p, const float* perm, float* u_star, float* v_star) {
  int rc = check_grid(grid);
  if (rc) return rc;
  if ((rc = check_ptrs("explicit_predict", {u, v, u_star, v_star}))) return rc;
  if (u_star == u || v_star == v || u_star == v || v_star == u) {
    set_error("explicit_predict: outputs must not alias inputs");
    return JAXIB_ERR_INVALID;
  }
  return launch_explicit((cudaStream_t)stream, make_gridf(*grid), u, v, p, perm, u_star, v_star, true);
}

int jaxib_ib_interp(jaxib_stream_t stream, const jaxib_grid* grid, double off_x, double off_y, const float* field,
                    const float* xp, const float* yp, int32_t n, float* out, int32_t* ci, int32_t* cj) {
  int rc = check_grid(grid);
  if (rc) return rc;
  if (n < 0) { set_error("ib_interp: n < 0"); return JAXIB_ERR_INVALID; }
  if (n == 0) return JAXIB_OK;
  if ((rc = check_ptrs("ib_interp", {field, xp, yp, out}))) return rc;
  if ((ci == nullptr) != (cj == nullptr)) { set_error("ib_interp: ci and cj must both be given"); return JAXIB_ERR_INVALID; }
  return launch_ib_interp((cudaStream_t)stream, make_gridf(*grid), (float)off_x, (float)off_y, field, xp, yp, n, out,
                          ci, cj);
}

int jaxib_ib_spread(jaxib_stream_t stream, const jaxib_grid* grid, double off_x, double off_y, const float* F,
                    const float* xp, const float* yp, const float* dS, int32_t n, double scale, float* out,
                    void* scratch, size_t scratch_bytes) {
  int rc = check_grid(grid);
  if (rc) return rc;
  if (n < 0) { set_error("ib_spread: n < 0"); return JAXIB_ERR_INVALID; }
  if (n == 0) return JAXIB_OK;
  if ((rc = check_ptrs("ib_spread", {F, xp, yp, dS, out}))) return rc;
  if (!scratch || scratch_bytes < 16) { set_error("ib_spread: needs 16 bytes of scratch"); return JAXIB_ERR_SCRATCH; }
  return launch_ib_spread((cudaStream_t)stream, make_gridf(*grid), (float)off_x, (float)off_y, F, xp, yp, dS, n,
                          (float)scale, out, reinterpret_cast<int*>(scratch), scratch_bytes);
}

size_t jaxib_ib_spread_scratch_bytes(const jaxib_grid* grid, int32_t n) {
  if (check_grid(grid)) return 0;
  return ib_spread_scratch_bytes(make_gridf(*grid), n);
}

static int check_bodies(const jaxib_bodies* b) {
  if (!b) { set_error("bodies is null"); return JAXIB_ERR_INVALID; }
  if (b->n_bodies < 0 || b->n_markers < 0) { set_error("bodies: negative counts"); return JAXIB_ERR_INVALID; }
  if (b->n_markers > 0 && (!b->body_offset || !b->x0 || !b->y0)) {
    set_error("bodies: null marker arrays");
    return JAXIB_ERR_INVALID;
  }
  if (b->n_markers > 0 && !(b->max_radius > 0.0)) { set_error("bodies: max_radius must be positive"); return JAXIB_ERR_INVALID; }
  return JAXIB_OK;
}

int jaxib_ib_force(jaxib_stream_t stream, const jaxib_grid* grid, const jaxib_bodies* bodies, const float* body_state,
                   float* u_star, float* v_star, void* scratch, size_t scratch_bytes) {
  int rc = check_grid(grid);
  if (rc) return rc;
  if ((rc = check_bodies(bodies))) return rc;
  if ((rc = check_ptrs("ib_force", {body_state, u_star, v_star}))) return rc;
  GridF g = make_gridf(*grid);
  size_t need = (size_t)5 * g.batch * bodies->n_markers * sizeof(float) + 256;
  if (!scratch || scratch_bytes < need) {
    set_error("ib_force: scratch too small: have %zu, need %zu", scratch_bytes, need);
    return JAXIB_ERR_SCRATCH;
  }
  uintptr_t a = (reinterpret_cast<uintptr_t>(scratch) + 255) & ~uintptr_t(255);
  return launch_ib_force((cudaStream_t)stream, g, *bodies, body_state, u_star, v_star, reinterpret_cast<float*>(a));
}

int jaxib_ib_marker_force(jaxib_stream_t stream, const jaxib_grid* grid, const jaxib_bodies* bodies,
                          const float* body_state, const float* u_star, const float* v_star, float* markers) {
  int rc = check_grid(grid);
  if (rc) return rc;
  if ((rc = check_bodies(bodies))) return rc;
  if ((rc = check_ptrs("ib_marker_force", {body_state, u_star, v_star, markers}))) return rc;
  return launch_ib_marker_force((cudaStream_t)stream, make_gridf(*grid), *bodies, body_state, u_star, v_star, markers);
}

int jaxib_ib_spread_bodies(jaxib_stream_t stream, const jaxib_grid* grid, const jaxib_bodies* bodies,
                           const float* body_state, const float* markers, float* u_star, float* v_star) {
  int rc = check_grid(grid);
  if (rc) return rc;
  if ((rc = check_bodies(bodies))) return rc;
  if ((rc = check_ptrs("ib_spread_bodies", {body_state, markers, u_star, v_star}))) return rc;
  return launch_ib_spread_bodies((cudaStream_t)stream, make_gridf(*grid), *bodies, body_state, markers, u_star, v_star);
}

static int check_pitch(const char* what, const jaxib_plan* plan, int32_t sp) {
  const GridF& g = plan_grid(plan);
  if (sp != 0 && (sp < g.ny / 2 + 4 || sp % 4 != 0)) {
    set_error("%s: spectrum pitch %d must be 0 or a multiple of 4 that is >= ny/2 + 4 = %d", what, sp, g.ny / 2 + 4);
    return JAXIB_ERR_INVALID;
  }
  return JAXIB_OK;
}

int jaxib_poisson_rows_fwd(jaxib_stream_t stream, const jaxib_plan* plan, const float* u, const float* v,
                           float* spectrum, int32_t sp) {
  if (!plan) { set_error("poisson_rows_fwd: plan is null"); return JAXIB_ERR_INVALID; }
  int rc = check_ptrs("poisson_rows_fwd", {u, v, spectrum});
  if (rc) return rc;
  if ((rc = check_pitch("poisson_rows_fwd", plan, sp))) return rc;
  return launch_rows_fwd((cudaStream_t)stream, plan, u, v, spectrum, sp);
}

int jaxib_poisson_cols(jaxib_stream_t stream, const jaxib_plan* plan, float* spectrum, int32_t sp, int32_t ncols,
                       int32_t col0) {
  if (!plan) { set_error("poisson_cols: plan is null"); return JAXIB_ERR_INVALID; }
  int rc = check_ptrs("poisson_cols", {spectrum});
  if (rc) return rc;
  const GridF& g = plan_grid(plan);
  if (sp < 0 || ncols < 0 || col0 < 0 || (sp > 0 && (sp % 2 != 0 || ncols > sp)) ||
      col0 + ncols > g.ny / 2 + 4 + 1024) {
    set_error("poisson_cols: bad column range (sp=%d ncols=%d col0=%d)", sp, ncols, col0);
    return JAXIB_ERR_INVALID;
  }
  return launch_cols((cudaStream_t)stream, plan, spectrum, sp, ncols, col0);
}

int jaxib_poisson_rows_inv(jaxib_stream_t stream, const jaxib_plan* plan, const float* spectrum, int32_t sp,
                           const float* u, const float* v, const float* p, float* u_out, float* v_out, float* p_out) {
  if (!plan) { set_error("poisson_rows_inv: plan is null"); return JAXIB_ERR_INVALID; }
  int rc = check_ptrs("poisson_rows_inv", {spectrum, u, v, p, u_out, v_out, p_out});
  if (rc) return rc;
  if ((rc = check_pitch("poisson_rows_inv", plan, sp))) return rc;
  return launch_rows_inv((cudaStream_t)stream, plan, spectrum, sp, u, v, p, u_out, v_out, p_out, nullptr);
}

int jaxib_pressure_solve(jaxib_stream_t stream, const jaxib_plan* plan, const float* u, const float* v, float* q,
                         void* scratch, size_t scratch_bytes) {
  if (!plan) { set_error("pressure_solve: plan is null"); return JAXIB_ERR_INVALID; }
  int rc = check_ptrs("pressure_solve", {u, v, q});
  if (rc) return rc;
  Scratch sc;
  if ((rc = carve(plan, 0, scratch, scratch_bytes, &sc))) return rc;
  return launch_project((cudaStream_t)stream, plan, u, v, nullptr, nullptr, nullptr, nullptr, q, sc.spectrum);
}

int jaxib_project(jaxib_stream_t stream, const jaxib_plan* plan, const float* u, const float* v, const float* p,
                  float* u_out, float* v_out, float* p_out, void* scratch, size_t scratch_bytes) {
  if (!plan) { set_error("project: plan is null"); return JAXIB_ERR_INVALID; }
  int rc = check_ptrs("project", {u, v, p, u_out, v_out, p_out});
  if (rc) return rc;
  Scratch sc;
  if ((rc = carve(plan, 0, scratch, scratch_bytes, &sc))) return rc;
  return launch_project((cudaStream_t)stream, plan, u, v, p, u_out, v_out, p_out, nullptr, sc.spectrum);
}

int jaxib_step(jaxib_stream_t stream, const jaxib_plan* plan, const jaxib_bodies* bodies, const float* body_table,
               const float* body_table_host, const float* perm, const double* wall_table, int32_t incremental_pressure,
               int32_t nsteps, float* u, float* v, float* p, void* scratch, size_t scratch_bytes) {
  if (!plan) { set_error("step: plan is null"); return JAXIB_ERR_INVALID; }
  int rc = check_ptrs("step", {u, v, p});
  if (rc) return rc;
  if (nsteps < 0) { set_error("step: nsteps < 0"); return JAXIB_ERR_INVALID; }
  const bool has_ib = bodies && bodies->n_markers > 0;
  if (has_ib) {
    if ((rc = check_bodies(bodies))) return rc;
    if (!body_table) { set_error("step: bodies given without body_table"); return JAXIB_ERR_INVALID; }
  }
  const GridF& g = plan_grid(plan);
  Scratch sc;
  if ((rc = carve(plan, has_ib ? bodies->n_markers : 0, scratch, scratch_bytes, &sc))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t state_stride = has_ib ? (size_t)g.batch * bodies->n_bodies * JAXIB_BODY_STATE_STRIDE : 0;
  // K4 early start (RowDeps, common.cuh): needs the ring predictor's completion counter and the team row kernel
  // Off by default: bit-identical (tests/test_gpu_round2.py), but measured on B200 at ib2048 85.8 -> 87.5 us per step (the
  // row kernel's CTAs only become resident once every CTA of the spreading kernel has started, which is after the
  // predictor has drained), ib1024 59.0 -> 55.4 us (profiles/r2n_early_start.log).  JAXIB_K4_EARLY=1 turns it on.
  static const bool k4_early_on = getenv("JAXIB_K4_EARLY") && atoi(getenv("JAXIB_K4_EARLY")) == 1;
  const bool k4_early = k4_early_on && has_ib && g.batch == 1 && plan_team_rows_fwd(plan) && explicit_ring_eligible(g);
  unsigned k1_total = 0;
  if (k4_early) cudaMemsetAsync(sc.sync, 0, sizeof(unsigned), s);
  for (int n = 0; n < nsteps; ++n) {
    // boundaries.py:519-542 update_BC: step n sees the wall values of its own time stamp
    GridF gn = g;
    if (wall_table) {
      const double* w = wall_table + JAXIB_WALL_STRIDE * (size_t)n;
      gn.uw_lo = (float)w[0]; gn.uw_hi = (float)w[1]; gn.vw_lo = (float)w[2]; gn.vw_hi = (float)w[3];
      gn.uwx_lo = (float)w[4]; gn.uwx_hi = (float)w[5]; gn.vwx_lo = (float)w[6]; gn.vwx_hi = (float)w[7];
    }
    const GridF* gov = wall_table ? &gn : nullptr;
    const float* pin = incremental_pressure ? p : nullptr;
    // IB chain on the side stream behind a split predictor (see OverlapCtx)
    RowBands near, rest;
    OverlapCtx* oc = nullptr;
    if (has_ib && body_table_host && g.batch == 1 && overlap_enabled() && explicit_ring_eligible(gn) &&
        body_row_bands(gn, *bodies, body_table_host + n * state_stride, &near, &rest))
      oc = overlap_ctx();
    bool split = false;
    static const bool trace = getenv("JAXIB_OVERLAP_TRACE") != nullptr;  // diagnostic: event time stamps of the fork / join
    cudaEvent_t tev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (oc) {
      int rc1 = JAXIB_OK;
      if (trace) { for (auto& e : tev) cudaEventCreate(&e); cudaEventRecord(tev[0], s); }
      split = launch_explicit_ring(s, gn, u, v, pin, perm, sc.u_star, sc.v_star, true, &rc1, &near);  // K1a
      if (split) {
        if (rc1) return rc1;
        cudaEventRecord(oc->fork, s);
        if (trace) cudaEventRecord(tev[1], s);
        cudaStreamWaitEvent(oc->side, oc->fork, 0);
        if (trace) cudaEventRecord(tev[2], oc->side);
        if ((rc = launch_ib_force(oc->side, gn, *bodies, body_table + n * state_stride, sc.u_star, sc.v_star, sc.markers)))
          return rc;
        cudaEventRecord(oc->join, oc->side);
        if (trace) cudaEventRecord(tev[3], oc->side);
        if (!launch_explicit_ring(s, gn, u, v, pin, perm, sc.u_star, sc.v_star, true, &rc1, &rest) || rc1) {  // K1b
          set_error("step: split predictor launch failed");
          return rc1 ? rc1 : JAXIB_ERR_CUDA;
        }
        if (trace) cudaEventRecord(tev[4], s);
        cudaStreamWaitEvent(s, oc->join, 0);
        if (trace) {
          cudaStreamSynchronize(s);
          float a, b, c, d;
          cudaEventElapsedTime(&a, tev[0], tev[1]); cudaEventElapsedTime(&b, tev[0], tev[2]);
          cudaEventElapsedTime(&c, tev[0], tev[3]); cudaEventElapsedTime(&d, tev[0], tev[4]);
          fprintf(stderr, "overlap trace (us from step start): K1a done %.1f | side stream starts %.1f | IB chain done %.1f | K1b done %.1f\n",
                  a * 1e3, b * 1e3, c * 1e3, d * 1e3);
          for (auto& e : tev) cudaEventDestroy(e);
        }
      }
    }
    RowDeps deps;
    memset(&deps, 0, sizeof(deps));
    if (!split && k4_early) {
      int rc1 = JAXIB_OK;
      unsigned inc = 0;
      if (launch_explicit_ring(s, gn, u, v, pin, perm, sc.u_star, sc.v_star, true, &rc1, nullptr, sc.sync, &inc)) {
        if (rc1) return rc1;
        k1_total += inc;
        deps.k1_counter = sc.sync; deps.k1_target = k1_total;
        deps.body_state = body_table + n * state_stride; deps.n_bodies = bodies->n_bodies;
        deps.half_extent = ib_half_extent(gn, *bodies);
        if ((rc = launch_ib_force(s, gn, *bodies, body_table + n * state_stride, sc.u_star, sc.v_star, sc.markers))) return rc;
        split = true;  // predictor and IB stage are enqueued
      }
    }
    if (!split) {
      // time_stepping.py:191-208
      if ((rc = launch_explicit(s, gn, u, v, pin, perm, sc.u_star, sc.v_star, true))) return rc;
      // time_stepping.py:210-223
      if (has_ib && (rc = launch_ib_force(s, gn, *bodies, body_table + n * state_stride, sc.u_star, sc.v_star, sc.markers)))
        return rc;
    }
    // time_stepping.py:246
    if ((rc = launch_project(s, plan, sc.u_star, sc.v_star, p, u, v, p, nullptr, sc.spectrum, nullptr, gov,
                             deps.k1_counter ? &deps : nullptr)))
      return rc;
  }
  return JAXIB_OK;
}

int jaxib_step_profile(jaxib_stream_t stream, const jaxib_plan* plan, const jaxib_bodies* bodies,
                       const float* body_table, const float* perm, int32_t incremental_pressure, int32_t nsteps,
                       float* u, float* v, float* p, void* scratch, size_t scratch_bytes, void* flush,
                       size_t flush_bytes, double* ms_out) {
  if (!plan || !ms_out) { set_error("step_profile: plan / ms_out is null"); return JAXIB_ERR_INVALID; }
  int rc = check_ptrs("step_profile", {u, v, p});
  if (rc) return rc;
  if (nsteps <= 0 || nsteps > 100000) { set_error("step_profile: nsteps out of range"); return JAXIB_ERR_INVALID; }
  const bool has_ib = bodies && bodies->n_markers > 0;
  if (has_ib) {
    if ((rc = check_bodies(bodies))) return rc;
    if (!body_table) { set_error("step_profile: bodies given without body_table"); return JAXIB_ERR_INVALID; }
  }
  const GridF& g = plan_grid(plan);
  Scratch sc;
  if ((rc = carve(plan, has_ib ? bodies->n_markers : 0, scratch, scratch_bytes, &sc))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t state_stride = has_ib ? (size_t)g.batch * bodies->n_bodies * JAXIB_BODY_STATE_STRIDE : 0;
  const int per = JAXIB_N_STAGES;  // start + 6 stage ends
  std::vector<cudaEvent_t> ev((size_t)per * nsteps);
  for (auto& e : ev) cudaEventCreate(&e);
  for (int n = 0; n < nsteps && !rc; ++n) {
    cudaEvent_t* e = ev.data() + (size_t)n * per;
    if (flush && flush_bytes) cudaMemsetAsync(flush, n & 0xff, flush_bytes, s);
    cudaEventRecord(e[0], s);
    rc = launch_explicit(s, g, u, v, incremental_pressure ? p : nullptr, perm, sc.u_star, sc.v_star, true);
    cudaEventRecord(e[1], s);
    StageHook hib{e + 2};
    if (!rc && has_ib) rc = launch_ib_force(s, g, *bodies, body_table + n * state_stride, sc.u_star, sc.v_star, sc.markers, &hib);
    if (!has_ib) { cudaEventRecord(e[2], s); cudaEventRecord(e[3], s); }
    StageHook hp{e + 4};
    if (!rc) rc = launch_project(s, plan, sc.u_star, sc.v_star, p, u, v, p, nullptr, sc.spectrum, &hp);
  }
  cudaError_t ce = cudaStreamSynchronize(s);
  for (int k = 0; k < JAXIB_N_STAGES; ++k) ms_out[k] = 0.0;
  if (!rc && ce == cudaSuccess) {
    for (int n = 0; n < nsteps; ++n) {
      cudaEvent_t* e = ev.data() + (size_t)n * per;
      float ms = 0.f;
      for (int k = 0; k < 6; ++k) {
        cudaEventElapsedTime(&ms, e[k], e[k + 1]);
        ms_out[k] += ms;
      }
      cudaEventElapsedTime(&ms, e[0], e[6]);
      ms_out[6] += ms;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  if (ce != cudaSuccess) { set_error("step_profile: %s", cudaGetErrorString(ce)); return JAXIB_ERR_CUDA; }
  return rc;
}

}  // extern "C"
