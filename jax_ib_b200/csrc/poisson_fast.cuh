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

}  // namespace jaxib
