// Internal declarations shared by the jax_ib_b200 CUDA translation units (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/jaxib_b200.h"

namespace jaxib {

// float32 view of jaxib_grid with the roundings the reference performs (JAX weak typing:
// Python-float scalars are rounded to float32 where they meet an array).
struct GridF {
  int nx, ny, batch;
  float lower_x, lower_y;
  float dx, dy;          // float32(step)
  float inv_dx, inv_dy;  // 1 / float32(step)
  float dt;
  float nu;              // float32(viscosity / density); diffusion disabled when has_nu == 0
  int has_nu;
  float dt_over_dx, dt_over_dy;  // float32(dt / step) computed in double (advection.py:273)
  float sx, sy, ssum;    // finite_differences.py:131: square(1 / float32(step)), and their sum
  float fx, fy;          // float32(force) / float32(density)
  float inv_rho;
  int bc, convect;
  float uw_lo, uw_hi, vw_lo, vw_hi;      // y walls
  float uwx_lo, uwx_hi, vwx_lo, vwx_hi;  // x walls (far field)
  int R;                 // IB window radius
  int row0, own_lo, own_hi, nx_global;  // slab decomposition (see jaxib_grid)
};

inline GridF make_gridf(const jaxib_grid& g) {
  GridF f;
  f.nx = g.nx; f.ny = g.ny; f.batch = g.batch < 1 ? 1 : g.batch;
  f.lower_x = (float)g.lower_x; f.lower_y = (float)g.lower_y;
  f.dx = (float)g.dx; f.dy = (float)g.dy;
  f.inv_dx = 1.0f / f.dx; f.inv_dy = 1.0f / f.dy;
  f.dt = (float)g.dt;
  f.has_nu = g.nu >= 0.0;
  f.nu = f.has_nu ? (float)g.nu : 0.0f;
  f.dt_over_dx = (float)(g.dt / g.dx); f.dt_over_dy = (float)(g.dt / g.dy);
  f.sx = f.inv_dx * f.inv_dx; f.sy = f.inv_dy * f.inv_dy; f.ssum = f.sx + f.sy;
  float rho = g.density > 0.0 ? (float)g.density : 1.0f;
  f.inv_rho = 1.0f / rho;
  f.fx = (float)g.force[0] / rho; f.fy = (float)g.force[1] / rho;
  f.bc = g.bc; f.convect = g.convect;
  f.uw_lo = (float)g.u_wall[0]; f.uw_hi = (float)g.u_wall[1];
  f.vw_lo = (float)g.v_wall[0]; f.vw_hi = (float)g.v_wall[1];
  f.uwx_lo = (float)g.u_wall_x[0]; f.uwx_hi = (float)g.u_wall_x[1];
  f.vwx_lo = (float)g.v_wall_x[0]; f.vwx_hi = (float)g.v_wall_x[1];
  f.R = g.ib_window > 0 ? g.ib_window : 6;
  f.row0 = g.row0; f.own_lo = g.own_lo; f.own_hi = g.own_hi; f.nx_global = g.nx_global;
  return f;
}

// ---- peer-memory slab mode (slab.cu): kernels store their results straight into the buffers of
// the slab that consumes them (NVLink stores through peer mappings) instead of handing them to a
// separate exchange.  cpr / world == 0 selects the ordinary single-GPU addressing.
constexpr int kMaxPeers = JAXIB_MAX_PEERS;

struct SpecScatter {          // K4 output: column block d of every spectrum row -> slab d's column buffer
  float2* base[kMaxPeers];    // [nx_global][cpr] complex on every slab
  int cpr;                    // columns per slab (0: plain [nx][sp] spectrum)
  unsigned magic;             // ceil(2^32 / cpr): k / cpr == umulhi(k, magic) for the k used here
  int row_offset;             // global row of this slab's first owned row
};

struct HaloOut {              // K6 output: owned boundary rows are also stored into the neighbours' halos
  float* lo[3];               // previous slab: its upper halo rows of u, v, p (row i of my block -> + i * ny)
  float* hi[3];               // next slab: its lower halo rows (row i -> + (i - (nx - halo)) * ny)
  int halo;                   // 0: off
};

__device__ __forceinline__ float2* spec_addr(const SpecScatter& S, int i, int k) {
  const unsigned d = __umulhi((unsigned)k, S.magic);
  return S.base[d] + (size_t)(S.row_offset + i) * S.cpr + (k - (int)d * S.cpr);
}

void set_error(const char* fmt, ...);
int check_grid(const jaxib_grid* g);
int check_launch(const char* what);

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// The kernels of a step form a dependency chain on one stream.  Launched with the programmatic
// stream-serialisation attribute, kernel N+1 may be scheduled while the last wave of kernel N is
// still running: its CTAs run their prologue (index arithmetic, twiddle tables into registers /
// shared memory -- nothing kernel N writes) and then block in pdl_wait() until kernel N has
// completed and flushed.  Every kernel calls pdl_trigger() first thing, and touches no global data
// of the chain (and writes nothing at all) before pdl_wait().  JAXIB_NO_PDL=1 restores plain launches.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled(int kernel_bit);  // JAXIB_PDL_MASK (default all) selects the kernels launched this way

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int kernel_bit, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled(kernel_bit) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- FFT plan (poisson.cu) ----------------------------------------------------------------
constexpr int kMaxStages = 16;
struct FftPlan {
  int n;                   // complex length
  int nstages;
  int radix[kMaxStages];   // DIF order: first pass acts on the whole length
};

// stage launchers (each only enqueues on `stream`)
int launch_explicit(cudaStream_t s, const GridF& g, const float* u, const float* v, const float* p,
                    const float* perm, float* out_u, float* out_v, bool predictor);
// optional per-stage event recording (jaxib_step_profile): ev[0..] are recorded after each kernel
struct StageHook {
  cudaEvent_t* ev;
};
int launch_ib_force(cudaStream_t s, const GridF& g, const jaxib_bodies& b, const float* body_state,
                    float* u_star, float* v_star, float* marker_scratch, const StageHook* hook = nullptr);
// Optional restriction of the two fused IB kernels to a range of bodies and of their markers (slab decomposition:
// the bodies near a slab, a host hint; counts of 0 = everything)
struct BodyRange {
  int body_first, body_count, marker_first, marker_count;
};
int launch_ib_marker_force(cudaStream_t s, const GridF& g, const jaxib_bodies& b, const float* body_state,
                           const float* u_star, const float* v_star, float* markers, const BodyRange* range = nullptr);
int launch_ib_spread_bodies(cudaStream_t s, const GridF& g, const jaxib_bodies& b, const float* body_state,
                            const float* markers, float* u_star, float* v_star, const BodyRange* range = nullptr);
bool launch_explicit_fast(cudaStream_t s, const GridF& g, const float* u, const float* v, const float* p,
                          const float* perm, float* out_u, float* out_v, bool predictor, int* rc);
// Row ranges of one predictor launch (stencil_ring.cu): the rows near the bodies are computed first so that the
// immersed-boundary kernels can run on a side stream while the rest of the grid is predicted (capi.cu).
constexpr int kMaxBands = 16;
struct RowBands {
  int n;                       // number of disjoint ranges, sorted
  int start[kMaxBands];
  int rows[kMaxBands];
  int chunk0[kMaxBands + 1];   // filled by the launcher: first CTA row-chunk of each range
};
bool explicit_ring_eligible(const GridF& g);
// done_counter (device, may be null): every consumer warp adds 1 when its rows are stored and fenced;
// *done_increment receives the number of such arrivals of this launch.
bool launch_explicit_ring(cudaStream_t s, const GridF& g, const float* u, const float* v, const float* p,
                          const float* perm, float* out_u, float* out_v, bool predictor, int* rc,
                          const RowBands* bands = nullptr, unsigned* done_counter = nullptr,
                          unsigned* done_increment = nullptr);

// Early start of the forward row transform (K4).  Its rows far from every body depend on the predictor (K1) only,
// not on the two latency-bound IB kernels that sit between K1 and K4 in the stream: a far row waits for K1's
// completion counter instead of the programmatic dependency on K3, so most of K4 runs UNDER the IB kernels.
// Forward progress: K4's CTAs can only start after every CTA of K3, hence of K2, hence of K1 has started (each
// kernel triggers its dependents in its prologue), so K1 is resident or finished whenever a K4 team spins; the
// spin is bounded and falls back to the ordinary dependency wait.
struct RowDeps {
  const unsigned* k1_counter;  // null: off
  unsigned k1_target;
  const float* body_state;     // [n_bodies][8] of this step (batch = 1)
  int n_bodies;
  int half_extent;             // body box half edge in cells (as the spreading kernel's)
};
int launch_ib_interp(cudaStream_t s, const GridF& g, float off_x, float off_y, const float* field,
                     const float* xp, const float* yp, int n, float* out, int32_t* ci, int32_t* cj);
int launch_ib_spread(cudaStream_t s, const GridF& g, float off_x, float off_y, const float* F,
                     const float* xp, const float* yp, const float* dS, int n, float scale, float* out,
                     int* box_scratch, size_t scratch_bytes);
size_t ib_spread_scratch_bytes(const GridF& g, int n);
int launch_project(cudaStream_t s, const jaxib_plan* plan, const float* u, const float* v, const float* p,
                   float* u_out, float* v_out, float* p_out, float* q_only, float* spectrum,
                   const StageHook* hook = nullptr, const GridF* gov = nullptr, const RowDeps* deps = nullptr);
// Poisson stages (sp = spectrum row pitch in complex values, 0 = ny/2 + 4)
int launch_rows_fwd(cudaStream_t s, const jaxib_plan* plan, const float* u, const float* v, float* spectrum, int sp,
                    const SpecScatter* sc = nullptr, const GridF* gov = nullptr, const RowDeps* deps = nullptr);
bool plan_team_rows_fwd(const jaxib_plan* plan);
int ib_half_extent(const GridF& g, const jaxib_bodies& b);
int launch_cols(cudaStream_t s, const jaxib_plan* plan, float* spectrum, int sp, int ncols, int col0,
                int col_begin = 0);
void plan_cols_geometry(const jaxib_plan* plan, int* cols_per_cta, int* ctas_per_sm);
int launch_rows_inv(cudaStream_t s, const jaxib_plan* plan, const float* spectrum, int sp, const float* u,
                    const float* v, const float* p, float* u_out, float* v_out, float* p_out, float* q_only,
                    const HaloOut* ho = nullptr, const GridF* gov = nullptr);
bool plan_fast_rows(const jaxib_plan* plan);
bool plan_fast_cols(const jaxib_plan* plan);
size_t spectrum_floats(const jaxib_plan* plan);
bool plan_sweep(const jaxib_plan* plan);
bool plan_col_sweep(const jaxib_plan* plan);
int plan_col_sweep_pitch(const jaxib_plan* plan);
int launch_cols_sweep_slab(cudaStream_t s, const jaxib_plan* plan, float* spectrum, int sp, int world, int rank,
                           float* const* tables, int mode);
size_t cols_sweep_slab_table_bytes(int world, int ncols_pad, int nx_global);
const GridF& plan_grid(const jaxib_plan* plan);

}  // namespace jaxib
