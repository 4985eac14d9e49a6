// Interface of the power-of-two fast Poisson kernels (poisson_fast.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace jaxib {

struct FastColParams {
  int nx, sp, ncols;
  const float2* tw;         // W_nx^k
  const float* lam_x_pos;   // float32(lx[freq(p)]) in digit-reversed position order
  const float* lam_y;       // float32(ly_l), padded to sp entries
  const float2* tw_team;    // team kernel: pre-laid-out twiddle tables [TA | TB] (team_twiddle_table)
  int first_wave;           // CTAs that may start early under PDL (set by the launcher)
  int limit;                // columns addressable from `spec` (row pitch minus the column offset)
};

struct FastRowParams {
  GridF g;
  const float2* tw;  // W_ny^k
  int sp;            // spectrum row pitch (complex values)
  int rows_per_group;
  float scale;       // 1 / (nx * ny/2)
  SpecScatter sc;    // rows_fwd, slab mode: spectrum columns go to the column owners
  HaloOut ho;        // rows_inv, slab mode: boundary rows also go to the neighbours' halos
  RowDeps deps;      // rows_fwd team kernel: early start of the rows far from every body (k1_counter == null: off)
};

bool fast_size_supported(int n);
bool fast_cols_supported(int n);  // the column kernel also handles 8192
void fast_radices(int n, int* radix4);  // radix4[3] == 0: three passes
int fast_rows_ctas_per_sm(int ny, bool inverse);
void fast_cols_geometry(int nx, int* cols_per_cta, int* ctas_per_sm);
int launch_cols_fast(cudaStream_t s, const FastColParams& P, int batch, float2* spec);
// second-generation column kernel (poisson_team.cu)
bool team_cols_supported(int n);
void team_radices(int n, int* radix4);
void team_twiddle_table(int n, std::vector<float2>* out);
int launch_cols_team(cudaStream_t s, const FastColParams& P, int batch, float2* spec);
int launch_rows_fwd_fast(cudaStream_t s, const FastRowParams& P, const float* u, const float* v, float2* spec);
int launch_rows_inv_fast(cudaStream_t s, const FastRowParams& P, const float2* spec, const float* u, const float* v,
                         const float* p, float* uo, float* vo, float* po, bool project);

// second-generation row kernels (poisson_team_rows.cu)
bool team_rows_supported(int n);  // n = ny / 2
void team_row_twiddle_table(int n, std::vector<float2>* out);
int launch_rows_fwd_team(cudaStream_t s, const FastRowParams& P, const float2* tables, const float* u, const float* v,
                         float2* spec);
int launch_rows_inv_team(cudaStream_t s, const FastRowParams& P, const float2* tables, const float2* spec, const float* u,
                         const float* v, const float* p, float* uo, float* vo, float* po, bool project);

// third generation of the periodic solve (poisson_team_rows.cu, "x-sweep"): row FFTs + two first-order sweeps along x
struct SweepParams {
  GridF g;
  int L, C, img;          // rows per chunk, chunks per grid, float2 entries of a row image
  const float* rho;       // [img] rho of the slot's wavenumber; < 0: padding hole; 0: the k = 0 column
  const float* lg;        // [img] log2(rho), finite
  const float* gain;      // [img] -dx^2 rho / (ny/2); k = 0: 1 / (ny/2)
  const double* rho_d;    // [img]
  float2* a;              // [batch][nx][img]  chunk-local forward sweep
  float2* agg;            // [batch][C][img]   sum_r rho^r a[r] of each chunk
  float2* cA;             // [batch][C][img]   forward sweep value entering the chunk
  float2* cB;             // [batch][C][img]   backward sweep value entering the chunk (= b of the next chunk's first row)
};

struct SweepGeom {
  int L, C, img;      // rows per chunk, chunks, float2 entries per row image
  int ntb;            // teams (rows in flight) per CTA
  int team_threads, max_teams;
};
bool sweep_geometry(int nx, int ny, int batch, SweepGeom* out);  // false: row length not supported
void sweep_tables(int ny, double dx, double dy, std::vector<float>* rho, std::vector<float>* lg, std::vector<float>* gain,
                  std::vector<double>* rho_d);
int launch_rows_fwd_sweep(cudaStream_t s, const SweepParams& P, int ntb, const float2* tables, const float* u, const float* v);
int launch_sweep_carry(cudaStream_t s, const SweepParams& P);
int launch_rows_inv_sweep(cudaStream_t s, const SweepParams& P, int ntb, const float2* tables, const float* u, const float* v,
                          const float* p, float* uo, float* vo, float* po, bool project);

// K5 third generation (poisson_sweep.cu): the periodic x-direction solved by two first-order sweeps, no transform
struct ColSweepParams {
  int nx, sp, ncols;    // rows, spectrum row pitch (complex), stored complex columns
  int segs;             // ceil(nx / 16): segments of rows per column
  const float2* cst;    // [7][cst_pitch] per-column constants (cols_sweep_tables)
  int cst_pitch;
  double dx2;           // dx^2
  double out_scale;     // nx of the whole column (see cols_sweep_tables)
  // slab decomposition (modes 1, 2): this GPU holds rows [rank nx, (rank + 1) nx) of nx_global
  int world, rank, nx_global, ncols_pad;
  int use_tma;          // set by the launcher: the strip is moved by tensor copies (cp.async.bulk.tensor.2d)
  float* slab_table[kMaxPeers];  // every slab's table of block maps + column 0 (cols_sweep_slab_table_bytes)
};
bool cols_sweep_supported(int nx);
void cols_sweep_tables(int nx, int nx_global, int ncols, double dx, const double* ly_re, const double* ly_im,
                       std::vector<float2>* out, int* pitch);
int launch_cols_sweep(cudaStream_t s, const ColSweepParams& P, int batch, float2* spec, int mode = 0);
size_t cols_sweep_slab_table_bytes(int world, int ncols_pad, int nx_global);

}  // namespace jaxib
