/*
 * jaxib_b200.h -- C ABI of the B200-native jax_ib per-timestep hot path.
 *
 * This is the drop-in boundary: every entry point replaces one pluggable callable (or a fused
 * group of them) of hashimmg/jax_IB's step function.  Citations are to files under the reference
 * tree (jax_ib/base/...).  The reference is pure Python/JAX and has no FFI of its own; the
 * binding a maintainer would add is an XLA-FFI custom call per entry point (see INTEGRATION.md,
 * jax_ib_b200/csrc/xla_ffi_shim.cc) or the ctypes stub in jax_ib_b200/_lib.py.
 *
 * Conventions
 *   - All array pointers are DEVICE pointers to float32, C-order [batch][nx][ny] (axis 1 = y is
 *     contiguous), exactly the reference's GridArray.data layout (grids.py:718 indexing='ij').
 *     Staggering (grids.py:632-637): u at offset (1, .5), v at (.5, 1), p at (.5, .5).
 *   - Every call only ENQUEUES work on `stream` (a cudaStream_t); none synchronises.
 *   - Return value: 0 on success, negative jaxib_status otherwise; jaxib_last_error() gives text.
 *   - All simulation state lives in the caller's buffers or in a jaxib_plan.  Process-wide, created on
 *     first use and never part of a result: one internal side stream (+ events) per device for the
 *     overlapped IB chain (jaxib_step) and for the slab decomposition's overlapped push (jaxib_slab_step),
 *     the once-per-function shared-memory opt-ins, and the environment switches read once (JAXIB_*).
 *     One host thread per device may drive the library at a time (the XLA runtime's model as well).
 *   - Inputs are const; in-place use is allowed only where stated.
 */
#ifndef JAXIB_B200_H_
#define JAXIB_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* jaxib_stream_t; /* cudaStream_t */

typedef enum {
  JAXIB_OK = 0,
  JAXIB_ERR_INVALID = -1,     /* bad argument (null pointer, non-positive size, ...) */
  JAXIB_ERR_UNSUPPORTED = -2, /* size / option outside what the kernels implement */
  JAXIB_ERR_SCRATCH = -3,     /* scratch buffer too small */
  JAXIB_ERR_CUDA = -4         /* CUDA runtime error (launch, allocation) */
} jaxib_status;

/* Boundary-condition families (boundaries.py:852-879 periodic; :636-663 moving wall / channel). */
typedef enum {
  JAXIB_BC_PERIODIC = 0, /* periodic x and y; pressure periodic                                  */
  JAXIB_BC_CHANNEL = 1,  /* periodic x; Dirichlet velocity / homogeneous-Neumann pressure in y   */
  JAXIB_BC_FAR_FIELD = 2 /* Dirichlet velocity on all four walls (boundaries.py:666-693), homogeneous-
                            Neumann pressure on both axes: pressure.py:134-154 solve_fast_diag_Far_Field */
} jaxib_bc;

/* Advection schemes (advection.py:120-124, :206-280, :325-333). */
typedef enum {
  JAXIB_CONVECT_UPWIND = 0,
  JAXIB_CONVECT_VAN_LEER = 1,
  JAXIB_CONVECT_VAN_LEER_LIMITERS = 2,
  JAXIB_CONVECT_NONE = 3
} jaxib_convect;

/* Grid + physics of one step.  Scalars are doubles because the reference holds them as Python
 * floats and rounds them to float32 only where they meet an array (e.g. float32(dt/dx), not
 * float32(dt)/float32(dx)); the library does the same rounding internally. */
typedef struct {
  int32_t nx, ny;      /* cells per axis (grids.py:563-624)                                      */
  int32_t batch;       /* independent simulations stacked on the leading axis (>= 1)            */
  double lower_x, lower_y; /* domain lower corner: mesh point i lies at lower + (i + offset) * step (grids.py:665) and
                            * the marker windows are centred on floor((x - lower) / step - offset)              */
  double dx, dy;       /* grid step = (upper - lower) / n                                       */
  double dt;           /* time step                                                             */
  double nu;           /* viscosity / density (equations.py:73); < 0 disables diffusion        */
  double density;      /* forcing is divided by it (equations.py:75)                            */
  int32_t bc;          /* jaxib_bc                                                              */
  int32_t convect;     /* jaxib_convect                                                         */
  double u_wall[2];    /* channel: u at the lower / upper y wall (bc_values[1] of velocity[0])  */
  double v_wall[2];    /* channel: v at the lower / upper y wall                                */
  double force[2];     /* constant body force (forcing(v) of the Taylor demo), per component    */
  int32_t ib_window;   /* Gaussian truncation radius R in cells (window 2R x 2R), 1..16; default 6.  The spreading
                        * Gaussian has width dx in BOTH directions (IBM_Force.py:67): with dy < dx pass
                        * ceil(R * dx / dy) to keep R standard deviations along y (the Python API does)    */
  /* Slab decomposition along x (all zero for a full grid): the arrays hold global rows
   * [row0, row0 + nx) (halo rows included).  jaxib_ib_marker_force computes forces only for markers
   * whose u-cell row floor(xp/dx - 1) lies in [own_lo, own_hi) and leaves the entries of the others
   * untouched (own_hi <= own_lo: all markers); spreading writes every local row.  nx_global > 0
   * marks a slab for the Poisson row stages (jaxib_poisson_rows_fwd / _rows_inv): the nx local rows
   * are "open" -- row -1 of u and spectrum row nx are read from the memory next to the arrays (the
   * caller's halo) instead of wrapping -- and transforms are normalised with nx_global.            */
  int32_t row0;
  int32_t own_lo, own_hi;
  int32_t nx_global;
  double u_wall_x[2];  /* far field: u at the lower / upper x wall (bc_values[0] of velocity[0])       */
  double v_wall_x[2];  /* far field: v at the lower / upper x wall                                     */
} jaxib_grid;

/* Rigid bodies of ONE simulation (identical marker topology across the batch).
 * x0/y0: body-frame marker coordinates returned by the reference's shape_fn (IBM_Force.py:29);
 * body_offset[b]..body_offset[b+1] is body b's (closed-curve) marker range. */
typedef struct {
  int32_t n_bodies;
  int32_t n_markers;
  const int32_t* body_offset; /* device [n_bodies + 1]                                         */
  const float* x0;            /* device [n_markers]                                             */
  const float* y0;            /* device [n_markers]                                             */
  double max_radius;          /* max over markers of hypot(x0, y0): bounds every body's extent  */
  /* Optional (may be NULL): device [n_markers][4] floats per marker k of body b = (x0, y0 of the NEXT
   * marker on b's closed curve -- roll(-1), IBM_Force.py:59-63 --, b as int32 bits, 0).  It removes the
   * marker -> body search and one dependent memory round trip from the marker-force kernel.          */
  const float* marker_pack;
  /* Optional (may be NULL): caller-owned device scratch of tile_flags_len int32, ZERO-INITIALISED once.  The
   * marker-force kernel marks the 8 x 8-cell tiles (of every body's bounding box, per velocity component) that
   * some marker window touches; the spreading kernel skips unmarked tiles at once and clears the marks it
   * consumes, so the buffer is all zero again after every step.  Needs batch * n_bodies * 2 * T^2 entries,
   * T = ceil((2 * (ceil(max_radius / min(dx, dy)) + ib_window + 2) + 1) / 8); too short = not used.  One
   * buffer per concurrently stepping grid (slab rank / stepper).                                        */
  int32_t* tile_flags;
  int32_t tile_flags_len;
} jaxib_bodies;

/* Per-step, per-body kinematic state, evaluated by the HOST from the user's Displacement_EQ /
 * Rotation_EQ callables (kinematics.py; jacrev at IBM_Force.py:101-104): 8 floats per body
 *   [cos(theta), sin(theta), xc, yc, U0x, U0y, Omega, 0]
 * laid out [step][batch][n_bodies][8]. */
#define JAXIB_BODY_STATE_STRIDE 8
#define JAXIB_WALL_STRIDE 8
#define JAXIB_MAX_PEERS 16

typedef struct jaxib_plan jaxib_plan; /* opaque: FFT twiddles, Laplacian eigenvalues, radix plan */

const char* jaxib_version(void);
const char* jaxib_last_error(void);
/* Compute capability the kernels were compiled for (100 => sm_100a). */
int jaxib_compiled_arch(void);

/* ---- plan ------------------------------------------------------------------------------- */
/* Replaces the host-side set-up of pressure.py:94-130 (array_utils.laplacian_matrix :168-182 +
 * jax_cfd fast_diagonalization.pseudoinverse): eigenvalues in float64, inverse cutoff 10*eps32. */
int jaxib_plan_create(const jaxib_grid* grid, jaxib_plan** out);
int jaxib_plan_destroy(jaxib_plan* plan);
/* Bytes of device scratch jaxib_step / jaxib_project / jaxib_ib_force need for this plan
 * (n_markers = total markers per simulation, 0 if no bodies). */
size_t jaxib_scratch_bytes(const jaxib_plan* plan, int32_t n_markers);

/* ---- stage: explicit terms + predictor ---------------------------------------------------- */
/* equations.py:42-83 navier_stokes_explicit_terms: (du, dv) = convect(v) + nu*laplacian(v) +
 * force/density [- perm*v/density when perm != NULL: penalty/util_funs.py:6-16].             */
int jaxib_explicit_terms(jaxib_stream_t stream, const jaxib_grid* grid, const float* u, const float* v,
                         const float* perm, float* du, float* dv);
/* time_stepping.py:191-208: u* = u + dt*F(u) - forward_difference(p) (p == NULL: no gradient,
 * the penalty stepper time_stepping.py:345).  Outputs must not alias inputs.                  */
int jaxib_explicit_predict(jaxib_stream_t stream, const jaxib_grid* grid, const float* u, const float* v,
                           const float* p, const float* perm, float* u_star, float* v_star);

/* ---- stage: immersed boundary -------------------------------------------------------------- */
/* convolution_functions.py:11-37 new_surf_fn (the `surface_fn` plug-in): Gaussian E->L
 * interpolation of `field` (offset = (off_x, off_y)) at n points; also returns the owning cell
 * floor(xp/dx - off_x), floor(yp/dy - off_y) (IBM_Force.py:35) when ci/cj != NULL.  batch = 1.  */
int jaxib_ib_interp(jaxib_stream_t stream, const jaxib_grid* grid, double off_x, double off_y,
                    const float* field, const float* xp, const float* yp, int32_t n, float* out,
                    int32_t* ci, int32_t* cj);
/* IBM_Force.py:66-87: out[i,j] += scale * sum_k F_k * delta(r_k; dx) * dS_k.  Cells outside every
 * marker window are left untouched (zero `out` first to obtain the bare force field).  batch = 1.
 * Scratch (device, 4-byte aligned): with jaxib_ib_spread_scratch_bytes(grid, n) bytes the markers are
 * counting-sorted into bins of 32 x 32 cells by their owning cell and every 8 x 8 tile gathers, in marker
 * order, from the bins its window reaches: O(n x window) work (the north star's "markers sorted by cell,
 * deterministic gather, no float atomics").  With less scratch (at least 16 bytes: the bounding box)
 * every tile inside the bounding box culls all n points instead.  Same result either way up to the
 * summation order of float32 (marker order in both; the tile's partial sums are grouped differently). */
size_t jaxib_ib_spread_scratch_bytes(const jaxib_grid* grid, int32_t n);
int jaxib_ib_spread(jaxib_stream_t stream, const jaxib_grid* grid, double off_x, double off_y,
                    const float* F, const float* xp, const float* yp, const float* dS, int32_t n,
                    double scale, float* out, void* scratch, size_t scratch_bytes);
/* IBM_Force.py:109-115 calc_IBM_force_NEW_MULTIPLE fused with time_stepping.py:223:
 * u** = u* + dt * f_IB, in place on (u_star, v_star).  body_state: [batch][n_bodies][8].
 * Marker arrays (xp, yp, dS, Fx, Fy: 5 * batch * n_markers floats) are left in `scratch`.     */
int jaxib_ib_force(jaxib_stream_t stream, const jaxib_grid* grid, const jaxib_bodies* bodies,
                   const float* body_state, float* u_star, float* v_star, void* scratch, size_t scratch_bytes);

/* ---- stage: pressure projection ------------------------------------------------------------ */
/* pressure.py:94-130 solve_fast_diag / solve_fast_diag_moving_wall (the `pressure_solve`
 * plug-in): q = pinv(Laplacian)(divergence(u, v)).                                            */
int jaxib_pressure_solve(jaxib_stream_t stream, const jaxib_plan* plan, const float* u, const float* v,
                         float* q, void* scratch, size_t scratch_bytes);
/* pressure.py:58-91 projection_and_update_pressure: q = solve; p_out = q + p; (u,v)_out =
 * (u,v) - forward_difference(q) [+ impose_bc for channel].  Outputs may alias the inputs.     */
int jaxib_project(jaxib_stream_t stream, const jaxib_plan* plan, const float* u, const float* v, const float* p,
                  float* u_out, float* v_out, float* p_out, void* scratch, size_t scratch_bytes);

/* ---- stage-level entry points for the slab decomposition -------------------------------------- */
/* New capability (the reference only shards MARKERS with jax.pmap, grid replicated:
 * IBM_Force.py:77-87, convolution_functions.py:28-35; no grid decomposition exists there).  A
 * slab rank calls the stages of one step separately and places its own exchange between them
 * (jax_ib_b200/slab.py): the spectrum is transposed (all-to-all) on either side of the column
 * transform and halo rows are refreshed after the projection.  Marker forces need no exchange:
 * with 2 * ib_window + 4 halo rows every slab computes the forces of the markers near it itself.
 *
 * jaxib_ib_force split in two.  `markers`: [5][batch][n_markers] floats (xp, yp, dS, Fx, Fy);
 * Fx, Fy are written only for the markers selected by grid->own_lo / own_hi.                   */
int jaxib_ib_marker_force(jaxib_stream_t stream, const jaxib_grid* grid, const jaxib_bodies* bodies,
                          const float* body_state, const float* u_star, const float* v_star, float* markers);
int jaxib_ib_spread_bodies(jaxib_stream_t stream, const jaxib_grid* grid, const jaxib_bodies* bodies,
                           const float* body_state, const float* markers, float* u_star, float* v_star);
/* jaxib_project split in three.  `spectrum` is [plan nx (+1 in slab mode for rows_inv)][sp] complex
 * float32; sp = row pitch in complex values (0 = ny/2 + 4; otherwise a multiple of 4, >= ny/2 + 4).
 * rows_fwd: divergence of (u, v) + real row FFT.  cols: in-place column FFT, division by the
 * eigenvalues (pressure.py:94-130), inverse column FFT for `ncols` columns holding the y-wavenumbers
 * col0 .. col0+ncols-1 (the plan's nx is the GLOBAL row count).  rows_inv: inverse row FFT fused
 * with p_out = p + q, (u, v)_out = (u, v) - grad q; it reads spectrum row i+1 for row i.        */
int jaxib_poisson_rows_fwd(jaxib_stream_t stream, const jaxib_plan* plan, const float* u, const float* v,
                           float* spectrum, int32_t sp);
int jaxib_poisson_cols(jaxib_stream_t stream, const jaxib_plan* plan, float* spectrum, int32_t sp, int32_t ncols,
                       int32_t col0);
int jaxib_poisson_rows_inv(jaxib_stream_t stream, const jaxib_plan* plan, const float* spectrum, int32_t sp,
                           const float* u, const float* v, const float* p, float* u_out, float* v_out, float* p_out);

/* ---- slab-decomposed step over peer-mapped memory ---------------------------------------------- */
/* One simulation on `world` GPUs (one process each), x split into slabs of nloc = nx / world rows.
 * The buffers below exist on every slab at the same sizes and are mapped into every process (CUDA
 * VMM / IPC peer mappings made by the host side; jax_ib_b200/slab.py uses torch symmetric memory);
 * entry [d] is slab d's buffer as addressed from THIS process, entry [rank] is the local one.
 * The kernels of a step store straight into the buffer of the slab that consumes the data (NVLink
 * stores) with a flag barrier after each exchanging phase -- see jax_ib_b200/csrc/slab.cu:
 *   default      the x-direction of the pressure solve is the column sweep (csrc/poisson_sweep.cu) run slab by
 *                slab: every slab stores its two carried-value maps per column and part + its rows of column 0
 *                into the head of every peer's colbuf, and its boundary rows into the neighbours' halos; 2
 *                barriers per step; ny/2 in {512..4096}, any nx divisible by world; results equal one GPU's
 *                to float32 round-off (rel-L2 <= 5e-6);
 *   JAXIB_SLAB_FFT=1 (environment)  round-1 path: both spectrum transposes through colbuf / specbuf, column FFT,
 *                3 barriers; additionally nx in {512..8192}; bit-identical to one GPU running the column FFT. */
typedef struct {
  int32_t world, rank;
  int32_t halo;                      /* halo rows on either side, in [2 * ib_window + 4, nloc]       */
  int32_t cpr;                       /* spectrum columns per slab: multiple of 8, world*cpr >= ny/2+4 */
  int32_t seam_idle;                 /* host hint: no marker window reaches the periodic seam in this block of
                                      * steps (rows < ib_window + 3 or >= nx - ib_window - 3): the first slab skips
                                      * the top-marker forces and the seam-row spread.  0 = always run them.      */
  int32_t body_first, body_count;    /* host hint (0 count = all): in this block of steps only the bodies
                                      * [body_first, body_first + body_count) and their markers                    */
  int32_t marker_first, marker_count;/* [marker_first, marker_first + marker_count) come near this slab's rows     */
  int32_t reserved_;
  float* u[JAXIB_MAX_PEERS];         /* [nloc + 2 halo][ny], owned rows in the middle                */
  float* v[JAXIB_MAX_PEERS];
  float* p[JAXIB_MAX_PEERS];
  float* markers[JAXIB_MAX_PEERS];   /* [5][n_markers]: xp, yp, dS, Fx, Fy (NULL without bodies)     */
  float* colbuf[JAXIB_MAX_PEERS];    /* [nx][cpr] complex, zero-initialised (sweep path: the slab tables)  */
  float* specbuf[JAXIB_MAX_PEERS];   /* [nloc + 1][world * cpr] complex                              */
  uint32_t* flags[JAXIB_MAX_PEERS];  /* 128 words, zero-initialised: barrier flags + error word [64] */
} jaxib_slab;
/* `nsteps` steps of slab `slab->rank`, in place on its u, v, p.  plan_rows: plan of the slab grid
 * (nx = nloc, nx_global = nx); plan_cols: plan of the global grid; grid: the GLOBAL grid.
 * body_table: [nsteps][n_bodies][8].  u_star, v_star: local scratch, [nloc + 2 halo][ny] each.
 * *epoch (host): barrier generation counter, starts at 0, advanced by nsteps; must be identical on all
 * slabs.  phase_lo..phase_hi (0..3) and barriers = 0 let a test run several slabs of ONE GPU in
 * lockstep from one process; production use is phases 0..3 with barriers = 1.                     */
int jaxib_slab_step(jaxib_stream_t stream, const jaxib_plan* plan_rows, const jaxib_plan* plan_cols,
                    const jaxib_grid* grid, const jaxib_slab* slab, const jaxib_bodies* bodies,
                    const float* body_table, int32_t nsteps, int32_t phase_lo, int32_t phase_hi, int32_t barriers,
                    uint32_t* epoch, float* u_star, float* v_star);

/* ---- fused step ------------------------------------------------------------------------------ */
/* time_stepping.py:144-255 step_fn, `nsteps` times, in place on (u, v, p):
 *   predict -> [IB interp / force / spread] -> project.
 * bodies == NULL skips the IB stage.  body_table: [nsteps][batch][n_bodies][8] (device).
 * body_table_host (may be NULL): the same table in HOST memory.  When given (batch = 1, upwind periodic
 * grids) the library knows where the bodies are and hides the two latency-bound IB kernels: the
 * predictor covers the rows near the bodies first, then the IB kernels run on an internal
 * high-priority side stream while the predictor covers the rest of the grid; the projection waits
 * for both.  Results are identical to the serial order.
 * perm (may be NULL): [batch][nx][ny] Brinkman permeability.  incremental_pressure == 0 selects
 * the penalty stepper (no -grad p_old in the predictor, time_stepping.py:345).
 * wall_table (HOST memory, may be NULL): [nsteps][JAXIB_WALL_STRIDE = 8] = u_lo, u_hi, v_lo, v_hi on the y walls
 * then u_lo, u_hi, v_lo, v_hi on the x walls (far field only), the wall values of a channel / far-field
 * grid at the time stamp of each step -- what update_BC (boundaries.py:519-542) re-evaluates
 * from the Moving_wall bc_fn every step (:636-663); NULL keeps the plan's constant u_wall / v_wall.
 * The clock / body centres are host-side state (boundaries.py:519-542, particle_motion.py:38-66)
 * and are not touched here.                                                                    */
int jaxib_step(jaxib_stream_t stream, const jaxib_plan* plan, const jaxib_bodies* bodies, const float* body_table,
               const float* body_table_host, const float* perm, const double* wall_table, int32_t incremental_pressure,
               int32_t nsteps, float* u, float* v, float* p, void* scratch, size_t scratch_bytes);

/* Measurement aid (bench.py): runs `nsteps` steps like jaxib_step but brackets every stage of every
 * step with CUDA events on `stream`, optionally overwrites `flush` (flush_bytes, > L2) before each
 * step outside the timed brackets, SYNCHRONISES the stream (the only entry point that does) and
 * returns the summed milliseconds per stage in ms_out[JAXIB_N_STAGES]:
 *   0 explicit_predict, 1 ib_interp_force, 2 ib_spread, 3 rows_fwd, 4 cols, 5 rows_inv, 6 whole step. */
#define JAXIB_N_STAGES 7
int jaxib_step_profile(jaxib_stream_t stream, const jaxib_plan* plan, const jaxib_bodies* bodies,
                       const float* body_table, const float* perm, int32_t incremental_pressure, int32_t nsteps,
                       float* u, float* v, float* p, void* scratch, size_t scratch_bytes, void* flush,
                       size_t flush_bytes, double* ms_out);

/* ---- tracers (MD/CFD_MD_coupling.py:7-27 + MD/simulate.py:81-97, kT = 0 or caller noise) ----- */
/* R[n][2] += (F_pair + [u,v](R)) * dt/(mass*gamma) + sqrt(2 kT dt/(mass gamma)) * xi, masked
 * by is_mobile (NULL = all mobile).  pair_force / xi may be NULL.                              */
int jaxib_tracer_step(jaxib_stream_t stream, const jaxib_grid* grid, const float* u, const float* v,
                      const float* pair_force, const float* xi, const uint8_t* is_mobile, double dt_md,
                      double mass_gamma, double kT, int32_t n, float* R);
/* Bilinear E->L sample of one field at n points (CFD_MD_coupling.py:7-19 interpolate_pbc).     */
int jaxib_bilinear_sample(jaxib_stream_t stream, const jaxib_grid* grid, double off_x, double off_y,
                          const float* field, const float* R, int32_t n, float* out);
/* Species-masked all-pairs harmonic/Morse force (MD/interaction_potential.py:6-22 through
 * jax_md smap.pair + quantity.force): h, r0 are [n_species][n_species] device matrices.        */
int jaxib_pair_force(jaxib_stream_t stream, const float* R, const int32_t* species, int32_t n, int32_t n_species,
                     const float* h, const float* r0, double D0, double alpha, double k, double box_x, double box_y,
                     float* F);

/* ---- penalty mask (penalty/util_funs.py:26-123) ----------------------------------------------- */
/* perm[i,j] = sum_b K/2 * (1 + tanh((R_b(theta)^2 - r^2)/width)); Rtheta: [n_bodies][ntheta],
 * centers: [n_bodies][2] (device).                                                              */
int jaxib_penalty_perm(jaxib_stream_t stream, const jaxib_grid* grid, const float* centers, const float* Rtheta,
                       int32_t n_bodies, int32_t ntheta, double K, double width, float* perm);

#ifdef __cplusplus
}
#endif
#endif /* JAXIB_B200_H_ */
