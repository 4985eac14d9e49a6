// K4/K5/K6: pressure projection by fast diagonalisation with a hand-written FFT.
//
// Reference: pressure.py:58-130 (projection_and_update_pressure, solve_fast_diag,
// solve_fast_diag_moving_wall) over jax_cfd.fast_diagonalization.pseudoinverse:
//     rhs = divergence(v);  q = irfftn( D * rfftn(rhs) );  D = 1/(lx_k + ly_l) if |.| > 10 eps32 else 0
//     p += q;  u -= (q[i+1,j]-q[i,j])/dx;  v -= (q[i,j+1]-q[i,j])/dy        (A4 of SURVEY.md)
// with lx_k = (2cos(2 pi k/Nx) - 2)/dx^2 (eigenvalues of array_utils.laplacian_matrix :168-173).
//
// Three kernels, 68 B/cell-step of algorithmic HBM traffic together with K1:
//   rows_fwd  (K4): divergence fused into a real FFT along y (contiguous axis).  A real row of Ny
//                   values is transformed as Ny/2 packed complex values and untangled; the real
//                   Nyquist coefficient is packed into the imaginary part of the DC coefficient, so
//                   the half spectrum is exactly [Nx][Ny/2] complex = one word per cell.
//   cols      (K5): complex FFT along x (strided axis) of a group of adjacent ky columns held in
//                   shared memory, multiply by D (computed in float64 from the two eigenvalue
//                   vectors, rounded once to float32 as the reference does), inverse FFT, in place.
//                   The packed DC/Nyquist column is separated through its Hermitian symmetry.
//   rows_inv  (K6): inverse real FFT along y fused with the gradient, the velocity correction
//                   and the pressure update.  Row i needs q[i+1], so a CTA that owns rows
//                   [i0, i1) transforms rows i0..i1 and finalises row i-1 when row i is ready.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "fft.cuh"
#include "poisson_fast.cuh"

struct jaxib_plan {
  jaxib_grid grid;
  jaxib::GridF gf;
  jaxib::FftPlan row;  // complex length ny/2
  jaxib::FftPlan col;  // complex length nx
  float2* tw_row;      // W_ny^k,  k < ny
  float2* tw_col;      // W_nx^k,  k < nx
  int* row_pos;        // position of frequency k after the forward row transform, k < ny/2
  int* col_pos;        // same for the column transform, k < nx
  double* lam_x_pos;   // lx[freq(p)] indexed by position p (digit-reversed), p < nx
  double* lam_y;       // ly_l, l <= ny/2
  int row_threads, col_threads, cols_per_cta, rows_per_cta;
  size_t row_smem, col_smem;
  // power-of-two fast paths (poisson_fast.cu)
  bool fast_rows, fast_cols;
  bool team_cols;               // K5 second generation (poisson_team.cu); implies fast_cols
  bool team_rows;               // K6 second generation (poisson_team_rows.cu); implies fast_rows
  bool team_rows_fwd;           // K4 second generation: row lengths ny <= 2048 (measured: the radix-16 first pass
                                // of longer rows spills at 64 registers and loses to the first generation)
  int fast_rows_per_group;      // K4
  int fast_rows_per_group_inv;  // K6 (each group also transforms the row after its last one)
  float* lam_x_pos_f;  // float32 copies of the eigenvalue tables (lam_y_f padded to ny/2 + 4)
  float* lam_y_f;
  float2* tw_dct;      // channel / far field: exp(-i pi k / (2 ny))
  float2* tw_dct_col;  // far field: exp(-i pi k / (2 nx))
  double* lam_x_nat;   // far field: Neumann eigenvalues (2cos(pi k/nx) - 2)/dx^2 in natural order k < nx
  float2* tw_team_col; // team column kernel: twiddle tables in shared-memory layout
  float2* tw_team_row; // team row kernels: ditto
  // third generation of the periodic solve (x-sweep, poisson_team_rows.cu): no column transform
  bool sweep;
  jaxib::SweepGeom sweep_geo;
  float* sw_rho;
  float* sw_lg;
  float* sw_gain;
  double* sw_rho_d;
  // K5 third generation (poisson_sweep.cu): x periodic (periodic and channel grids), any nx
  bool col_sweep;
  float2* cs_tab;
  int cs_pitch;
};

namespace jaxib {
namespace {

constexpr double kCutoff = 10.0 * 1.1920928955078125e-07;  // 10 * finfo(float32).eps
constexpr int kLamPad = 1024;  // eigenvalue tables are padded with 1.0: slab ranks hold padded column ranges

bool factor_plan(int n, FftPlan* pl) {
  pl->n = n;
  pl->nstages = 0;
  int e2 = 0, e3 = 0, e5 = 0;
  while (n % 2 == 0) { n /= 2; ++e2; }
  while (n % 3 == 0) { n /= 3; ++e3; }
  while (n % 5 == 0) { n /= 5; ++e5; }
  if (n != 1) return false;
  auto push = [&](int r) { pl->radix[pl->nstages++] = r; };
  // small odd radices first (they act on long, conflict-free strides), then 16/8/4/2
  for (int i = 0; i < e3; ++i) push(3);
  for (int i = 0; i < e5; ++i) push(5);
  while (e2 >= 7 || e2 == 4) { push(16); e2 -= 4; }
  while (e2 >= 3) { push(8); e2 -= 3; }
  if (e2 == 2) push(4);
  if (e2 == 1) push(2);
  return pl->nstages <= kMaxStages && pl->nstages > 0;
}

int pos_of_freq(const FftPlan& pl, int k) {
  int p = 0, len = pl.n;
  for (int s = 0; s < pl.nstages; ++s) {
    int r = pl.radix[s];
    len /= r;
    p += (k % r) * len;
    k /= r;
  }
  return p;
}

template <typename T>
int upload(T** dst, const std::vector<T>& src) {
  if (cudaMalloc((void**)dst, src.size() * sizeof(T)) != cudaSuccess) return JAXIB_ERR_CUDA;
  if (cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess)
    return JAXIB_ERR_CUDA;
  return JAXIB_OK;
}

struct RowParams {
  GridF g;
  FftPlan pl;
  const float2* tw;
  const int* pos;
  int rows_per_cta;
  int sp;       // spectrum row pitch in complex values (>= ny/2 + 1, multiple of 4)
  float scale;  // 1 / (nx * ny/2)
};

__device__ __forceinline__ int wrapi(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

// ---- K4 -------------------------------------------------------------------------------------------
__global__ void rows_fwd_kernel(RowParams P, const float* __restrict__ u, const float* __restrict__ v,
                                float2* __restrict__ spec) {
  extern __shared__ float2 buf[];
  const GridF& g = P.g;
  const int n2 = g.ny >> 1;
  const size_t plane = (size_t)g.nx * g.ny;
  const int b = blockIdx.y;
  u += b * plane; v += b * plane; spec += (size_t)b * g.nx * P.sp;
  const int i0 = blockIdx.x * P.rows_per_cta;
  const int i1 = min(i0 + P.rows_per_cta, g.nx);
  for (int i = i0; i < i1; ++i) {
    const float2* ur = reinterpret_cast<const float2*>(u + (size_t)i * g.ny);
    const float2* um = reinterpret_cast<const float2*>(u + (ptrdiff_t)(g.nx_global ? i - 1 : wrapi(i - 1, g.nx)) * g.ny);
    const float* vrow = v + (size_t)i * g.ny;
    const float2* vr = reinterpret_cast<const float2*>(vrow);
    for (int n = threadIdx.x; n < n2; n += blockDim.x) {
      float2 a = ur[n], c = um[n], d = vr[n];
      float vm;  // v[i][2n-1]
      if (n > 0) vm = vrow[2 * n - 1];
      else vm = g.bc == JAXIB_BC_PERIODIC ? vrow[g.ny - 1] : g.vw_lo;
      // finite_differences.py:139-146: backward differences, summed
      float r0 = (a.x - c.x) * g.inv_dx + (d.x - vm) * g.inv_dy;
      float r1 = (a.y - c.y) * g.inv_dx + (d.y - d.x) * g.inv_dy;
      buf[padi(n)] = make_float2(r0, r1);
    }
    __syncthreads();
    fft_forward(P.pl, buf, 1, 0, P.tw, 2);
    // untangle the packed real transform and store the half spectrum (DC.im carries Nyquist)
    float2* out = spec + (size_t)i * P.sp;
    for (int k = threadIdx.x; k <= n2 / 2; k += blockDim.x) {
      if (k == 0) {
        float2 z = buf[padi(P.pos[0])];
        // DC and Nyquist are both real; they are kept as separate columns: packing them into one
        // complex column would tie the noise floor of the (most amplified) DC column to the size
        // of the Nyquist column
        out[0] = make_float2(z.x + z.y, 0.0f);
        out[n2] = make_float2(z.x - z.y, 0.0f);
      } else {
        const int km = n2 - k;
        float2 zk = buf[padi(P.pos[k])], zm = cconj(buf[padi(P.pos[km])]);
        float2 e = cscale(cadd(zk, zm), 0.5f);
        float2 o = cscale(mul_mi(csub(zk, zm)), 0.5f);
        float2 t = cmul(__ldg(P.tw + k), o);
        out[k] = cadd(e, t);
        if (km != k) out[km] = cconj(csub(e, t));
      }
    }
    if (threadIdx.x < 3) out[n2 + 1 + threadIdx.x] = make_float2(0.0f, 0.0f);  // padding columns stay finite
    __syncthreads();
  }
}

// ---- K6 -------------------------------------------------------------------------------------------
template <bool PROJECT>
__global__ void rows_inv_kernel(RowParams P, const float2* __restrict__ spec, const float* u, const float* v,
                                const float* p, float* u_out, float* v_out, float* p_out) {
  extern __shared__ float2 smem[];
  const GridF& g = P.g;
  const int n2 = g.ny >> 1;
  const int pitch = padi(n2) + 1;
  const size_t plane = (size_t)g.nx * g.ny;
  const int b = blockIdx.y;
  spec += (size_t)b * g.nx * P.sp;
  if (PROJECT) { u += b * plane; v += b * plane; p += b * plane; u_out += b * plane; v_out += b * plane; }
  p_out += b * plane;
  const int i0 = blockIdx.x * P.rows_per_cta;
  const int i1 = min(i0 + P.rows_per_cta, g.nx);
  const int last = PROJECT ? i1 : i1 - 1;  // projection needs q of the row after the last owned one
  for (int ii = i0; ii <= last; ++ii) {
    const int i = (ii >= g.nx && g.nx_global == 0) ? ii - g.nx : ii;  // slab mode: spectrum row nx exists (halo)
    float2* cur = smem + ((ii - i0) & 1) * pitch;
    const float2* in = spec + (size_t)i * P.sp;
    for (int k = threadIdx.x; k <= n2 / 2; k += blockDim.x) {
      if (k == 0) {
        const float x0 = in[0].x, xn = in[n2].x;  // X[0], X[ny/2]: both real
        cur[padi(P.pos[0])] = make_float2(0.5f * (x0 + xn), 0.5f * (x0 - xn));
      } else {
        const int km = n2 - k;
        float2 xk = in[k], xm = cconj(in[km]);
        float2 e = cscale(cadd(xk, xm), 0.5f);
        float2 o = cmulc(cscale(csub(xk, xm), 0.5f), __ldg(P.tw + k));
        cur[padi(P.pos[k])] = cadd(e, mul_pi(o));
        if (km != k) cur[padi(P.pos[km])] = cadd(cconj(e), mul_pi(cconj(o)));
      }
    }
    __syncthreads();
    fft_inverse(P.pl, cur, 1, 0, P.tw, 2);
    if (!PROJECT) {
      float2* qo = reinterpret_cast<float2*>(p_out + (size_t)i * g.ny);
      for (int n = threadIdx.x; n < n2; n += blockDim.x) qo[n] = cscale(cur[padi(n)], P.scale);
    } else if (ii > i0) {
      const int im = ii - 1;  // owned row to finalise
      const float2* prv = smem + ((ii - 1 - i0) & 1) * pitch;
      const size_t ro = (size_t)im * g.ny;
      const float2* u2 = reinterpret_cast<const float2*>(u + ro);
      const float2* v2 = reinterpret_cast<const float2*>(v + ro);
      const float2* p2 = reinterpret_cast<const float2*>(p + ro);
      float2* uo = reinterpret_cast<float2*>(u_out + ro);
      float2* vo = reinterpret_cast<float2*>(v_out + ro);
      float2* po = reinterpret_cast<float2*>(p_out + ro);
      for (int n = threadIdx.x; n < n2; n += blockDim.x) {
        float2 q0 = cscale(prv[padi(n)], P.scale);
        float2 q1 = cscale(cur[padi(n)], P.scale);
        float qn;  // q[im][2n+2]
        if (n + 1 < n2) qn = prv[padi(n + 1)].x * P.scale;
        else qn = g.bc == JAXIB_BC_PERIODIC ? prv[padi(0)].x * P.scale : q0.y;  // Neumann ghost = edge
        float2 uu = u2[n], vv = v2[n], pp = p2[n];
        uu.x -= (q1.x - q0.x) * g.inv_dx;
        uu.y -= (q1.y - q0.y) * g.inv_dx;
        vv.x -= (q0.y - q0.x) * g.inv_dy;
        vv.y -= (qn - q0.y) * g.inv_dy;
        if (g.bc == JAXIB_BC_CHANNEL && n + 1 == n2) vv.y = g.vw_hi;  // impose_bc (boundaries.py:405-419)
        uo[n] = uu;
        vo[n] = vv;
        po[n] = make_float2(q0.x + pp.x, q0.y + pp.y);
      }
    }
    __syncthreads();
  }
}

// ---- channel rows: DCT-II / DCT-III along y (Neumann pressure) --------------------------------------
// pressure.py:111-130 diagonalises the Neumann Laplacian (array_utils.py:175-182) with eigh; its
// eigenvectors are the DCT-II basis, lambda_l = (2cos(pi l/Ny) - 2)/dy^2.  Makhoul's algorithm:
//   v[n] = x[2n], v[N-1-n] = x[2n+1];  V = FFT_N(v);  C[k] = Re(w_k V[k]),  w_k = exp(-i pi k / 2N)
// and back:  V[k] = conj(w_k) (C[k] - i C[N-k]),  v = IFFT_N(V).  The real coefficients are stored
// [Nx][Ny] (pitch 2*sp floats); the column kernel treats adjacent coefficient pairs as one complex
// column (packed_all).  The forward/inverse pair is exact, so no orthonormal scaling is needed.
struct DctParams {
  GridF g;
  FftPlan pl;          // complex length ny
  const float2* tw;    // W_ny^k
  const float2* twd;   // w_k = exp(-i pi k / (2 ny)), k < ny
  const int* pos;      // digit-reversed position of frequency k (length ny)
  int rows_per_cta;
  int sp;              // spectrum pitch in complex units (real pitch = 2 sp)
  float scale;         // 1 / (nx * ny)
};

__global__ void rows_fwd_dct_kernel(DctParams P, const float* __restrict__ u, const float* __restrict__ v,
                                    float* __restrict__ spec) {
  extern __shared__ float2 buf[];
  const GridF& g = P.g;
  const int N = g.ny;
  const size_t plane = (size_t)g.nx * N;
  u += blockIdx.y * plane; v += blockIdx.y * plane; spec += (size_t)blockIdx.y * g.nx * 2 * P.sp;
  const int i0 = blockIdx.x * P.rows_per_cta, i1 = min(i0 + P.rows_per_cta, g.nx);
  for (int i = i0; i < i1; ++i) {
    const float* ur = u + (size_t)i * N;
    const float* um = u + (size_t)wrapi(i - 1, g.nx) * N;
    const float* vr = v + (size_t)i * N;
    const bool xwall = g.bc == JAXIB_BC_FAR_FIELD && i == 0;  // u(-1, j) is the lower x wall value (edge-aligned u)
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      // divergence with the channel ghost rule: v(i,-1) is the lower wall value (boundaries.py:205-220)
      const float vm = j > 0 ? vr[j - 1] : g.vw_lo;
      const float r = (ur[j] - (xwall ? g.uwx_lo : um[j])) * g.inv_dx + (vr[j] - vm) * g.inv_dy;
      const int n = (j & 1) ? N - 1 - (j >> 1) : (j >> 1);
      buf[padi(n)] = make_float2(r, 0.0f);
    }
    __syncthreads();
    fft_forward(P.pl, buf, 1, 0, P.tw, 1);
    float* out = spec + (size_t)i * 2 * P.sp;
    for (int k = threadIdx.x; k < N; k += blockDim.x) out[k] = cmul(__ldg(P.twd + k), buf[padi(P.pos[k])]).x;
    __syncthreads();
  }
}

template <bool PROJECT>
__global__ void rows_inv_dct_kernel(DctParams P, const float* __restrict__ spec, const float* u, const float* v,
                                    const float* p, float* u_out, float* v_out, float* p_out) {
  extern __shared__ float2 smem[];
  const GridF& g = P.g;
  const int N = g.ny;
  const int pitch = padi(N) + 1;
  float2* buf = smem;
  float* qrow[2] = {reinterpret_cast<float*>(smem + pitch), reinterpret_cast<float*>(smem + pitch) + N};
  const size_t plane = (size_t)g.nx * N;
  spec += (size_t)blockIdx.y * g.nx * 2 * P.sp;
  if (PROJECT) { u += blockIdx.y * plane; v += blockIdx.y * plane; p += blockIdx.y * plane; u_out += blockIdx.y * plane; v_out += blockIdx.y * plane; }
  p_out += blockIdx.y * plane;
  const int i0 = blockIdx.x * P.rows_per_cta, i1 = min(i0 + P.rows_per_cta, g.nx);
  const int last = PROJECT ? i1 : i1 - 1;
  for (int ii = i0; ii <= last; ++ii) {
    const int i = ii >= g.nx ? ii - g.nx : ii;
    const float* in = spec + (size_t)i * 2 * P.sp;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
      const float ck = in[k], cm = k > 0 ? in[N - k] : 0.0f;
      buf[padi(P.pos[k])] = k > 0 ? cmulc(make_float2(ck, -cm), __ldg(P.twd + k)) : make_float2(ck, 0.0f);
    }
    __syncthreads();
    fft_inverse(P.pl, buf, 1, 0, P.tw, 1);
    float* cur = qrow[(ii - i0) & 1];
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      const int n = (j & 1) ? N - 1 - (j >> 1) : (j >> 1);
      cur[j] = buf[padi(n)].x * P.scale;
    }
    __syncthreads();
    if (!PROJECT) {
      float* qo = p_out + (size_t)i * N;
      for (int j = threadIdx.x; j < N; j += blockDim.x) qo[j] = cur[j];
    } else if (ii > i0) {
      const float* prv = qrow[(ii - 1 - i0) & 1];
      const size_t ro = (size_t)(ii - 1) * N;
      const bool xlast = g.bc == JAXIB_BC_FAR_FIELD && ii - 1 == g.nx - 1;  // q(nx) = q(nx-1) (Neumann); u(nx-1) on the wall
      for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const float q0 = prv[j], q1 = xlast ? prv[j] : cur[j];
        const float qn = j + 1 < N ? prv[j + 1] : q0;  // homogeneous Neumann ghost = edge (boundaries.py:247-269)
        float vv = v[ro + j] - (qn - q0) * g.inv_dy;
        if (j + 1 == N) vv = g.vw_hi;                  // impose_bc: the last v row lies on the wall (boundaries.py:405-419)
        if (xlast) u_out[ro + j] = g.uwx_hi;           // impose_bc on the upper x wall
        else
        u_out[ro + j] = u[ro + j] - (q1 - q0) * g.inv_dx;
        v_out[ro + j] = vv;
        p_out[ro + j] = q0 + p[ro + j];
      }
    }
    __syncthreads();
  }
}

// ---- K5 -------------------------------------------------------------------------------------------
struct ColParams {
  int nx, sp, ncols;      // spectrum is [nx][sp]; ncols = number of stored columns
  FftPlan pl;
  const float2* tw;
  const int* pos;
  const double* lam_x_pos;
  const double* lam_y;
  int cols_per_cta;
  int packed_all;         // channel: every stored column is a pair of real (DCT) columns
};

__device__ __forceinline__ float inv_eig(double lx, double ly) {
  double l = lx + ly;
  return fabs(l) > kCutoff ? (float)(1.0 / l) : 0.0f;
}

__global__ void cols_kernel(ColParams P, float2* __restrict__ spec) {
  extern __shared__ float2 buf[];
  const int nx = P.nx;
  const int pitch = padi(nx) + 1;
  spec += (size_t)blockIdx.y * nx * P.sp;
  const int c0 = blockIdx.x * P.cols_per_cta;
  const int nc = min(P.cols_per_cta, P.ncols - c0);
  const int total = nx * nc;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int i = e / nc, c = e - i * nc;
    buf[c * pitch + padi(i)] = spec[(size_t)i * P.sp + c0 + c];
  }
  __syncthreads();
  fft_forward(P.pl, buf, nc, pitch, P.tw, 1);
  for (int c = 0; c < nc; ++c) {
    float2* col = buf + c * pitch;
    const int ky = c0 + c;
    const bool packed = P.packed_all;
    if (!packed) {
      const double ly = P.lam_y[ky];
      for (int q = threadIdx.x; q < nx; q += blockDim.x) {
        float d = inv_eig(P.lam_x_pos[q], ly);
        col[padi(q)] = cscale(col[padi(q)], d);
      }
    } else {
      // column holds A + iB with A, B the transforms of two REAL sequences a (eigenvalue ly_a) and
      // b (ly_b): A[k] = (Z[k] + conj Z[-k])/2, B[k] = (Z[k] - conj Z[-k])/(2i)
      const double lya = P.lam_y[2 * ky];
      const double lyb = P.lam_y[2 * ky + 1];
      for (int k = threadIdx.x; k <= nx / 2; k += blockDim.x) {
        const int km = k == 0 ? 0 : nx - k;
        const int p1 = P.pos[k], p2 = P.pos[km];
        const double lx = P.lam_x_pos[p1];
        const float da = inv_eig(lx, lya), db = inv_eig(lx, lyb);
        float2 z1 = col[padi(p1)], z2 = cconj(col[padi(p2)]);
        float2 a = cscale(cadd(z1, z2), 0.5f * da);
        float2 bq = cscale(mul_mi(csub(z1, z2)), 0.5f * db);
        col[padi(p1)] = cadd(a, mul_pi(bq));
        if (p2 != p1) col[padi(p2)] = cadd(cconj(a), mul_pi(cconj(bq)));
      }
    }
  }
  __syncthreads();
  fft_inverse(P.pl, buf, nc, pitch, P.tw, 1);
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int i = e / nc, c = e - i * nc;
    spec[(size_t)i * P.sp + c0 + c] = buf[c * pitch + padi(i)];
  }
}

// ---- K5, far field: DCT-II along x of PAIRS of real coefficient columns ------------------------------
// pressure.py:134-154 diagonalises the Neumann x Neumann Laplacian (array_utils.py:254-298) with eigh + tensordot;
// the eigenbasis is the DCT-II on both axes.  The row kernels leave real DCT-y coefficients [nx][ny]; two adjacent
// real columns (a, b) are one complex column z = a + i b.  Makhoul along x: reorder v[n] = z[2n], v[N-1-n] = z[2n+1],
// one complex FFT, split into the transforms A, B of the two real sequences through Hermitian symmetry,
// C[k] = Re(w_k V[k]), C[N-k] = -Im(w_k V[k]); scale by 1/(lx_k + ly_l) (cut-off 10 eps32); and back:
// V[k] = conj(w_k)(C[k] - i C[N-k]), recombine, inverse FFT, undo the reorder.  Exact forward/inverse pair: the
// 1/(nx ny) normalisation stays in the inverse row kernel.
struct ColDctParams {
  int nx, sp, ncols;       // spectrum [nx][sp] complex = [nx][2 sp] real; ncols complex columns
  FftPlan pl;
  const float2* tw;        // W_nx^k
  const float2* twd;       // exp(-i pi k / (2 nx))
  const int* pos;
  const double* lam_x;     // natural order, k < nx
  const double* lam_y;     // ly_l, l < ny
  int cols_per_cta;
};

__global__ void cols_dct_kernel(ColDctParams P, float2* __restrict__ spec) {
  extern __shared__ float2 buf[];
  const int N = P.nx;
  const int pitch = padi(N) + 1;
  spec += (size_t)blockIdx.y * N * P.sp;
  const int c0 = blockIdx.x * P.cols_per_cta;
  const int nc = min(P.cols_per_cta, P.ncols - c0);
  const int total = N * nc;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int i = e / nc, c = e - i * nc;
    const int n = (i & 1) ? N - 1 - (i >> 1) : (i >> 1);
    buf[c * pitch + padi(n)] = spec[(size_t)i * P.sp + c0 + c];
  }
  __syncthreads();
  fft_forward(P.pl, buf, nc, pitch, P.tw, 1);
  for (int c = 0; c < nc; ++c) {
    float2* col = buf + c * pitch;
    const double lya = P.lam_y[2 * (c0 + c)], lyb = P.lam_y[2 * (c0 + c) + 1];
    for (int k = threadIdx.x; k <= N / 2; k += blockDim.x) {
      const int km = k == 0 ? 0 : N - k;
      const int p1 = P.pos[k], p2 = P.pos[km];
      const float2 z1 = col[padi(p1)], z2 = cconj(col[padi(p2)]);
      // transforms of the two real sequences at frequency k (at N - k they are the conjugates)
      const float2 A = cscale(cadd(z1, z2), 0.5f), B = cscale(mul_mi(csub(z1, z2)), 0.5f);
      const float2 w = __ldg(P.twd + k);
      const float2 wa = cmul(w, A), wb = cmul(w, B);
      // DCT coefficients C[k] = Re(w V[k]), C[N-k] = -Im(w V[k]); their eigenvalues differ
      const double lk = P.lam_x[k], lm = k == 0 ? 0.0 : P.lam_x[km];
      float ak = wa.x * inv_eig(lk, lya), bk = wb.x * inv_eig(lk, lyb);
      float am = k == 0 ? 0.0f : -wa.y * inv_eig(lm, lya), bm = k == 0 ? 0.0f : -wb.y * inv_eig(lm, lyb);
      if (2 * k == N) { am = 0.0f; bm = 0.0f; }  // C[N/2] is its own mirror: V'[N/2] = conj(w)(1 - i) C[N/2] below
      // back: V'[k] = conj(w_k)(C[k] - i C[N-k])
      float2 Va, Vb;
      if (2 * k == N) {
        Va = cmulc(make_float2(ak, -ak), w);
        Vb = cmulc(make_float2(bk, -bk), w);
      } else {
        Va = cmulc(make_float2(ak, -am), w);
        Vb = cmulc(make_float2(bk, -bm), w);
      }
      col[padi(p1)] = cadd(Va, mul_pi(Vb));                          // Z'[k]   = Va + i Vb
      if (p2 != p1) col[padi(p2)] = cadd(cconj(Va), mul_pi(cconj(Vb)));  // Z'[N-k] = conj(Va) + i conj(Vb)
    }
  }
  __syncthreads();
  fft_inverse(P.pl, buf, nc, pitch, P.tw, 1);
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int i = e / nc, c = e - i * nc;
    const int n = (i & 1) ? N - 1 - (i >> 1) : (i >> 1);
    spec[(size_t)i * P.sp + c0 + c] = buf[c * pitch + padi(n)];
  }
}

}  // namespace

const GridF& plan_grid(const jaxib_plan* plan) { return plan->gf; }
size_t spectrum_floats(const jaxib_plan* plan) {
  const size_t plain = (size_t)plan->gf.batch * plan->gf.nx * (plan->gf.ny + 8);  // [nx][ny/2 + 4] complex
  if (!plan->sweep) return plain;
  // x-sweep: row images [nx][img] + three [chunks][img] arrays of carried values, complex
  const SweepGeom& sg = plan->sweep_geo;
  const size_t sweep = 2 * (size_t)plan->gf.batch * sg.img * ((size_t)plan->gf.nx + 3 * (size_t)sg.C);
  return sweep > plain ? sweep : plain;
}
bool plan_sweep(const jaxib_plan* plan) { return plan->sweep; }
bool plan_col_sweep(const jaxib_plan* plan) { return plan->col_sweep; }
int plan_col_sweep_pitch(const jaxib_plan* plan) { return plan->cs_pitch; }

// passes A (mode 1) / B (mode 2) of the slab-decomposed column sweep on this slab's spectrum rows [nloc (+ 1)][sp]
int launch_cols_sweep_slab(cudaStream_t s, const jaxib_plan* pl, float* spectrum, int sp, int world, int rank,
                           float* const* tables, int mode) {
  const GridF& g = pl->gf;
  if (!pl->col_sweep || g.nx_global <= 0 || g.bc != JAXIB_BC_PERIODIC) { set_error("slab column sweep: plan is not a periodic slab plan"); return JAXIB_ERR_INVALID; }
  ColSweepParams P;
  memset(&P, 0, sizeof(P));
  P.nx = g.nx; P.sp = sp; P.ncols = g.ny / 2 + 1; P.segs = (g.nx + 15) / 16;
  P.cst = pl->cs_tab; P.cst_pitch = pl->cs_pitch;
  P.dx2 = (double)g.dx * (double)g.dx; P.out_scale = (double)g.nx_global;
  P.world = world; P.rank = rank; P.nx_global = g.nx_global; P.ncols_pad = pl->cs_pitch;
  for (int d = 0; d < world; ++d) P.slab_table[d] = tables[d];
  return launch_cols_sweep(s, P, 1, reinterpret_cast<float2*>(spectrum), mode);
}

namespace {

struct StageParams {
  RowParams rp;
  FastRowParams frp;
  DctParams dp;
  dim3 rgrid;
  bool channel;
};

// sp = spectrum row pitch in complex values (0 = the default ny/2 + 4)
StageParams stage_params(const jaxib_plan* pl, int sp, const GridF* gov = nullptr) {
  StageParams S;
  const GridF& g = gov ? *gov : pl->gf;  // gov: the plan's grid with this step's wall values (update_BC)
  const int n2 = g.ny / 2;
  const double nxg = g.nx_global > 0 ? g.nx_global : g.nx;  // slab: normalise with the global length
  S.channel = g.bc != JAXIB_BC_PERIODIC;
  S.rp.g = g; S.rp.pl = pl->row; S.rp.tw = pl->tw_row; S.rp.pos = pl->row_pos; S.rp.rows_per_cta = pl->rows_per_cta;
  S.rp.sp = sp > 0 ? sp : n2 + 4;
  S.rp.scale = (float)(1.0 / (nxg * (double)n2));
  S.rgrid = dim3((g.nx + pl->rows_per_cta - 1) / pl->rows_per_cta, g.batch);
  S.frp.g = g; S.frp.tw = pl->tw_row; S.frp.sp = S.rp.sp; S.frp.rows_per_group = pl->fast_rows_per_group;  // launch_rows_inv overrides with the K6 value
  S.frp.scale = S.rp.scale;
  memset(&S.frp.sc, 0, sizeof(S.frp.sc));
  memset(&S.frp.ho, 0, sizeof(S.frp.ho));
  memset(&S.frp.deps, 0, sizeof(S.frp.deps));
  if (S.channel) {
    S.dp.g = g; S.dp.pl = pl->row; S.dp.tw = pl->tw_row; S.dp.twd = pl->tw_dct; S.dp.pos = pl->row_pos;
    S.dp.rows_per_cta = pl->rows_per_cta; S.dp.sp = S.rp.sp;
    S.dp.scale = (float)(1.0 / (nxg * (double)g.ny));
  }
  return S;
}

}  // namespace

// K4: divergence of (u, v) + forward row transform -> spectrum [nx][sp]
bool plan_team_rows_fwd(const jaxib_plan* pl) { return pl->team_rows_fwd; }

int launch_rows_fwd(cudaStream_t s, const jaxib_plan* pl, const float* u, const float* v, float* spectrum, int sp,
                    const SpecScatter* sc, const GridF* gov, const RowDeps* deps) {
  StageParams S = stage_params(pl, sp, gov);
  if (deps && pl->team_rows_fwd) S.frp.deps = *deps;
  float2* spec = reinterpret_cast<float2*>(spectrum);
  if (sc && sc->cpr) {
    if (!pl->fast_rows) { set_error("peer-memory slab mode needs ny/2 in {512, 1024, 2048, 4096}"); return JAXIB_ERR_UNSUPPORTED; }
    S.frp.sc = *sc;
  }
  if (S.channel) {
    rows_fwd_dct_kernel<<<S.rgrid, pl->row_threads, pl->row_smem, s>>>(S.dp, u, v, spectrum);
    return check_launch("rows_fwd_dct_kernel");
  }
  if (pl->team_rows_fwd) return launch_rows_fwd_team(s, S.frp, pl->tw_team_row, u, v, spec);
  if (pl->fast_rows) return launch_rows_fwd_fast(s, S.frp, u, v, spec);
  rows_fwd_kernel<<<S.rgrid, pl->row_threads, pl->row_smem, s>>>(S.rp, u, v, spec);
  return check_launch("rows_fwd_kernel");
}

// K5: column transform, division by the eigenvalues, inverse column transform, in place, for the
// `ncols` stored columns whose y-wavenumbers start at col0 (a slab rank holds a column range)
bool plan_fast_rows(const jaxib_plan* pl) { return pl->fast_rows; }
bool plan_fast_cols(const jaxib_plan* pl) { return pl->fast_cols; }

void plan_cols_geometry(const jaxib_plan* pl, int* cols_per_cta, int* ctas_per_sm) {
  if (pl->team_cols) { *cols_per_cta = pl->gf.nx >= 4096 ? 4 : 8; *ctas_per_sm = 1; }
  else if (pl->fast_cols) fast_cols_geometry(pl->gf.nx, cols_per_cta, ctas_per_sm);
  else { *cols_per_cta = pl->cols_per_cta; *ctas_per_sm = 1; }
}

// col_begin: first column (of the `sp` stored per row) to transform; col0: y-wavenumber of stored column 0
int launch_cols(cudaStream_t s, const jaxib_plan* pl, float* spectrum, int sp, int ncols, int col0, int col_begin) {
  const GridF& g = pl->gf;
  const int n2 = g.ny / 2;
  const bool channel = g.bc != JAXIB_BC_PERIODIC;
  if (sp <= 0) sp = n2 + 4;
  if (pl->col_sweep && g.nx_global == 0 && ncols <= 0 && col0 == 0 && col_begin == 0) {  // whole grid on this GPU: sweeps instead of a transform
    ColSweepParams P;
    memset(&P, 0, sizeof(P));
    P.nx = g.nx; P.sp = sp; P.ncols = channel ? n2 : n2 + 1; P.segs = (g.nx + 15) / 16;
    P.cst = pl->cs_tab; P.cst_pitch = pl->cs_pitch;
    P.dx2 = (double)g.dx * (double)g.dx; P.out_scale = (double)g.nx;
    return launch_cols_sweep(s, P, g.batch, reinterpret_cast<float2*>(spectrum));
  }
  float2* spec = reinterpret_cast<float2*>(spectrum) + col_begin;
  col0 += col_begin;
  if (pl->fast_cols) {
    FastColParams fcp;
    fcp.limit = sp - col_begin;
    fcp.nx = g.nx; fcp.sp = sp; fcp.ncols = ncols > 0 ? ncols : n2 + 1; fcp.tw = pl->tw_col;
    fcp.lam_x_pos = pl->lam_x_pos_f; fcp.lam_y = pl->lam_y_f + col0; fcp.tw_team = pl->tw_team_col;
    return pl->team_cols ? launch_cols_team(s, fcp, g.batch, spec) : launch_cols_fast(s, fcp, g.batch, spec);
  }
  if (g.bc == JAXIB_BC_FAR_FIELD) {
    ColDctParams dp;
    dp.nx = g.nx; dp.sp = sp; dp.ncols = ncols > 0 ? ncols : n2; dp.pl = pl->col; dp.tw = pl->tw_col; dp.twd = pl->tw_dct_col;
    dp.pos = pl->col_pos; dp.lam_x = pl->lam_x_nat; dp.lam_y = pl->lam_y + 2 * col0; dp.cols_per_cta = pl->cols_per_cta;
    dim3 dgrid((dp.ncols + pl->cols_per_cta - 1) / pl->cols_per_cta, g.batch);
    cols_dct_kernel<<<dgrid, pl->col_threads, pl->col_smem, s>>>(dp, spec);
    return check_launch("cols_dct_kernel");
  }
  ColParams cp;
  // channel: ny real DCT coefficients per row = ny/2 packed column pairs; periodic: ny/2+1 complex columns
  cp.nx = g.nx; cp.sp = sp; cp.ncols = ncols > 0 ? ncols : (channel ? n2 : n2 + 1); cp.pl = pl->col; cp.tw = pl->tw_col;
  cp.pos = pl->col_pos;
  cp.lam_x_pos = pl->lam_x_pos; cp.lam_y = pl->lam_y + (channel ? 2 * col0 : col0); cp.cols_per_cta = pl->cols_per_cta;
  cp.packed_all = channel;
  dim3 cgrid((cp.ncols + pl->cols_per_cta - 1) / pl->cols_per_cta, g.batch);
  cols_kernel<<<cgrid, pl->col_threads, pl->col_smem, s>>>(cp, spec);
  return check_launch("cols_kernel");
}

// K6: inverse row transform (+ projection and pressure update unless q_only)
int launch_rows_inv(cudaStream_t s, const jaxib_plan* pl, const float* spectrum, int sp, const float* u, const float* v,
                    const float* p, float* u_out, float* v_out, float* p_out, float* q_only, const HaloOut* ho,
                    const GridF* gov) {
  StageParams S = stage_params(pl, sp, gov);
  if (ho && ho->halo) {
    if (!pl->fast_rows) { set_error("peer-memory slab mode needs ny/2 in {512, 1024, 2048, 4096}"); return JAXIB_ERR_UNSUPPORTED; }
    S.frp.ho = *ho;
  }
  const GridF& g = pl->gf;
  const float2* spec = reinterpret_cast<const float2*>(spectrum);
  if (S.channel) {
    const size_t smem = pl->row_smem + 2 * (size_t)g.ny * sizeof(float);
    if (q_only)
      rows_inv_dct_kernel<false><<<S.rgrid, pl->row_threads, smem, s>>>(S.dp, spectrum, nullptr, nullptr, nullptr,
                                                                        nullptr, nullptr, q_only);
    else
      rows_inv_dct_kernel<true><<<S.rgrid, pl->row_threads, smem, s>>>(S.dp, spectrum, u, v, p, u_out, v_out, p_out);
    return check_launch("rows_inv_dct_kernel");
  }
  if (pl->team_rows)
    return launch_rows_inv_team(s, S.frp, pl->tw_team_row, spec, u, v, p, u_out, v_out, q_only ? q_only : p_out,
                                q_only == nullptr);
  if (pl->fast_rows) {
    S.frp.rows_per_group = pl->fast_rows_per_group_inv;
    return launch_rows_inv_fast(s, S.frp, spec, u, v, p, u_out, v_out, q_only ? q_only : p_out, q_only == nullptr);
  }
  if (q_only)
    rows_inv_kernel<false><<<S.rgrid, pl->row_threads, 2 * pl->row_smem, s>>>(S.rp, spec, nullptr, nullptr, nullptr,
                                                                             nullptr, nullptr, q_only);
  else
    rows_inv_kernel<true><<<S.rgrid, pl->row_threads, 2 * pl->row_smem, s>>>(S.rp, spec, u, v, p, u_out, v_out, p_out);
  return check_launch("rows_inv_kernel");
}

int launch_project(cudaStream_t s, const jaxib_plan* pl, const float* u, const float* v, const float* p, float* u_out,
                   float* v_out, float* p_out, float* q_only, float* spectrum, const StageHook* hook, const GridF* gov,
                   const RowDeps* deps) {
  if (pl->sweep) {  // periodic grids with team row kernels: FFT along y, two recurrences along x, no column pass
    const SweepGeom& sg = pl->sweep_geo;
    SweepParams P;
    P.g = gov ? *gov : pl->gf;
    P.L = sg.L; P.C = sg.C; P.img = sg.img;
    P.rho = pl->sw_rho; P.lg = pl->sw_lg; P.gain = pl->sw_gain; P.rho_d = pl->sw_rho_d;
    const size_t rows = (size_t)P.g.batch * P.g.nx * sg.img, carr = (size_t)P.g.batch * sg.C * sg.img;
    P.a = reinterpret_cast<float2*>(spectrum);
    P.agg = P.a + rows;
    P.cA = P.agg + carr;
    P.cB = P.cA + carr;
    int rc = launch_rows_fwd_sweep(s, P, sg.ntb, pl->tw_team_row, u, v);
    if (rc) return rc;
    if (hook) cudaEventRecord(hook->ev[0], s);
    if ((rc = launch_sweep_carry(s, P))) return rc;
    if (hook) cudaEventRecord(hook->ev[1], s);
    rc = launch_rows_inv_sweep(s, P, sg.ntb, pl->tw_team_row, u, v, p, u_out, v_out, q_only ? q_only : p_out, q_only == nullptr);
    if (hook) cudaEventRecord(hook->ev[2], s);
    return rc;
  }
  int rc = launch_rows_fwd(s, pl, u, v, spectrum, 0, nullptr, gov, deps);
  if (rc) return rc;
  if (hook) cudaEventRecord(hook->ev[0], s);
  rc = launch_cols(s, pl, spectrum, 0, 0, 0);
  if (rc) return rc;
  if (hook) cudaEventRecord(hook->ev[1], s);
  rc = launch_rows_inv(s, pl, spectrum, 0, u, v, p, u_out, v_out, p_out, q_only, nullptr, gov);
  if (hook) cudaEventRecord(hook->ev[2], s);
  return rc;
}

}  // namespace jaxib

using namespace jaxib;

extern "C" int jaxib_plan_create(const jaxib_grid* grid, jaxib_plan** out) {
  if (!out) { set_error("plan_create: out is null"); return JAXIB_ERR_INVALID; }
  *out = nullptr;
  int rc = check_grid(grid);
  if (rc) return rc;
  if (grid->ny % 2) { set_error("ny=%d must be even for the packed real FFT", grid->ny); return JAXIB_ERR_UNSUPPORTED; }
  jaxib_plan* pl = new jaxib_plan();
  memset(pl, 0, sizeof(*pl));
  pl->grid = *grid;
  pl->gf = make_gridf(*grid);
  const int nx = grid->nx, ny = grid->ny, n2 = ny / 2;
  const bool channel = grid->bc != JAXIB_BC_PERIODIC;  // Neumann pressure in y: row DCT (channel and far field)
  const bool far = grid->bc == JAXIB_BC_FAR_FIELD;     // Neumann pressure in x as well: column DCT
  const int nrow = channel ? ny : n2;  // complex length of the row transform (DCT via a length-ny FFT)
  if (!factor_plan(nrow, &pl->row) || !factor_plan(nx, &pl->col)) {
    set_error("grid %dx%d: sizes must factor into 2, 3 and 5", nx, ny);
    delete pl;
    return JAXIB_ERR_UNSUPPORTED;
  }
  // power-of-two sizes use the register-resident kernels of poisson_fast.cu; the column plan must
  // then carry the radix triple those kernels are instantiated with (it defines the digit-reversed
  // positions of the eigenvalue table)
  const bool aligned = (ny % 4) == 0;
  pl->fast_rows = grid->bc == JAXIB_BC_PERIODIC && aligned && fast_size_supported(n2) && getenv("JAXIB_NO_FAST_FFT") == nullptr;
  pl->fast_cols = grid->bc == JAXIB_BC_PERIODIC && aligned && fast_cols_supported(nx) && getenv("JAXIB_NO_FAST_FFT") == nullptr;
  {
    const char* e = getenv("JAXIB_COLS");  // JAXIB_COLS=fast selects the first-generation column kernel (A/B runs)
    pl->team_cols = pl->fast_cols && team_cols_supported(nx) && !(e && !strcmp(e, "fast"));
  }
  {
    const char* e = getenv("JAXIB_ROWS");  // JAXIB_ROWS=fast selects the first-generation row kernels (A/B runs)
    pl->team_rows = pl->fast_rows && team_rows_supported(n2) && !(e && !strcmp(e, "fast"));
    const char* ef = getenv("JAXIB_ROWS_FWD");
    pl->team_rows_fwd = pl->team_rows && (ef ? !strcmp(ef, "team") : n2 <= 1024);
  }
  if (pl->fast_cols) {
    pl->col.n = nx;
    if (pl->team_cols) team_radices(nx, pl->col.radix);
    else fast_radices(nx, pl->col.radix);  // three radices, four for nx = 8192
    pl->col.nstages = pl->col.radix[3] ? 4 : 3;
  }
  const double two_pi = 6.283185307179586476925286766559;
  std::vector<float2> twr(ny), twc(nx);
  for (int k = 0; k < ny; ++k) twr[k] = make_float2((float)cos(two_pi * k / ny), (float)-sin(two_pi * k / ny));
  for (int k = 0; k < nx; ++k) twc[k] = make_float2((float)cos(two_pi * k / nx), (float)-sin(two_pi * k / nx));
  std::vector<int> rpos(nrow), cpos(nx);
  for (int k = 0; k < nrow; ++k) rpos[k] = pos_of_freq(pl->row, k);
  for (int k = 0; k < nx; ++k) cpos[k] = pos_of_freq(pl->col, k);
  // eigenvalues of the circulant 1-D Laplacians (array_utils.py:168-173), float64
  std::vector<double> lxp(nx), ly((channel ? ny : n2 + 1) + kLamPad, 1.0);
  for (int k = 0; k < nx; ++k) lxp[cpos[k]] = (2.0 * cos(two_pi * k / nx) - 2.0) / (grid->dx * grid->dx);
  if (channel) {  // eigenvalues of laplacian_matrix_neumann (array_utils.py:175-182): DCT-II basis
    for (int l = 0; l < ny; ++l) ly[l] = (2.0 * cos(0.5 * two_pi * l / ny) - 2.0) / (grid->dy * grid->dy);
    std::vector<float2> twd(ny);
    for (int k = 0; k < ny; ++k)
      twd[k] = make_float2((float)cos(0.25 * two_pi * k / ny), (float)-sin(0.25 * two_pi * k / ny));
    rc = upload(&pl->tw_dct, twd);
    if (!rc && far) {  // eigenvalues of the Neumann Laplacian in x too (array_utils.py:254-298), natural order
      std::vector<double> lxn(nx);
      std::vector<float2> twdc(nx);
      for (int k = 0; k < nx; ++k) {
        lxn[k] = (2.0 * cos(0.5 * two_pi * k / nx) - 2.0) / (grid->dx * grid->dx);
        twdc[k] = make_float2((float)cos(0.25 * two_pi * k / nx), (float)-sin(0.25 * two_pi * k / nx));
      }
      rc = upload(&pl->lam_x_nat, lxn);
      if (!rc) rc = upload(&pl->tw_dct_col, twdc);
    }
  } else {
    for (int l = 0; l <= n2; ++l) ly[l] = (2.0 * cos(two_pi * l / ny) - 2.0) / (grid->dy * grid->dy);
  }
  if (!rc && pl->team_cols) {
    std::vector<float2> tt;
    team_twiddle_table(nx, &tt);
    rc = upload(&pl->tw_team_col, tt);
  }
  if (!rc && pl->team_rows) {
    std::vector<float2> tt;
    team_row_twiddle_table(n2, &tt);
    rc = upload(&pl->tw_team_row, tt);
  }
  {
    const char* e = getenv("JAXIB_SWEEP");  // JAXIB_SWEEP=0: keep the FFT x FFT solve (A/B runs)
    pl->sweep = pl->team_rows && grid->nx_global == 0 && (e && atoi(e) == 1) &&
                sweep_geometry(nx, ny, pl->gf.batch, &pl->sweep_geo);
    if (!rc && pl->sweep) {
      std::vector<float> r, l, gn;
      std::vector<double> rd;
      sweep_tables(ny, (double)pl->gf.dx, (double)pl->gf.dy, &r, &l, &gn, &rd);
      rc = upload(&pl->sw_rho, r);
      if (!rc) rc = upload(&pl->sw_lg, l);
      if (!rc) rc = upload(&pl->sw_gain, gn);
      if (!rc) rc = upload(&pl->sw_rho_d, rd);
    }
  }
  {
    const char* e = getenv("JAXIB_COL_SWEEP");  // JAXIB_COL_SWEEP=0: keep the column FFT (A/B runs)
    pl->col_sweep = !far && cols_sweep_supported(nx) && !(e && atoi(e) == 0);  // nx_global > 0: the slab passes
    if (!rc && pl->col_sweep) {
      const int nc = channel ? n2 : n2 + 1;
      std::vector<double> lre(nc), lim(nc);
      for (int c = 0; c < nc; ++c) {
        lre[c] = channel ? ly[2 * c] : ly[c];
        lim[c] = channel ? ly[2 * c + 1] : ly[c];
      }
      std::vector<float2> tab;
      cols_sweep_tables(nx, grid->nx_global, nc, (double)pl->gf.dx, lre.data(), lim.data(), &tab, &pl->cs_pitch);
      rc = upload(&pl->cs_tab, tab);
    }
  }
  if (!rc) rc = upload(&pl->tw_row, twr);
  if (!rc) rc = upload(&pl->tw_col, twc);
  if (!rc) rc = upload(&pl->row_pos, rpos);
  if (!rc) rc = upload(&pl->col_pos, cpos);
  if (!rc) rc = upload(&pl->lam_x_pos, lxp);
  if (!rc) rc = upload(&pl->lam_y, ly);
  {
    std::vector<float> lxf(nx), lyf(n2 + 4 + kLamPad, 1.0f);
    for (int k = 0; k < nx; ++k) lxf[k] = (float)lxp[k];
    for (int l = 0; l <= n2 && !channel; ++l) lyf[l] = (float)ly[l];
    if (!rc) rc = upload(&pl->lam_x_pos_f, lxf);
    if (!rc) rc = upload(&pl->lam_y_f, lyf);
  }
  if (rc) { set_error("plan_create: device allocation failed"); jaxib_plan_destroy(pl); return rc; }

  int dev = 0, sms = 148, max_smem = 227 * 1024;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  auto pick_threads = [](int n) {
    int t = ((n / 8 + 31) / 32) * 32;
    return t < 64 ? 64 : (t > 256 ? 256 : t);
  };
  pl->row_threads = pick_threads(nrow);
  pl->row_smem = (size_t)(padi(nrow) + 1) * sizeof(float2);
  long long rows_total = (long long)nx * pl->gf.batch;
  int rpc = (int)(rows_total / (3LL * sms));
  pl->rows_per_cta = rpc < 1 ? 1 : (rpc > 16 ? 16 : rpc);
  {  // fast row kernels: one row group per CTA, `rpg` rows in sequence.  Wave-quantisation model, checked
     // against a sweep on B200 (tools/tune_rows.py): time ~ waves * (rpg + 1 [+ 1 for K6's extra row]),
     // waves = ceil(groups / resident CTAs).  E.g. a 1024-row slab of 8192 columns: 3 rows/CTA = 342 CTAs =
     // 3 waves (53 us) vs 7 rows/CTA = 147 CTAs = 1 wave (43 us).
    auto pick = [&](bool inverse) {
      const long long slots = (long long)sms * (pl->fast_rows ? fast_rows_ctas_per_sm(ny, inverse) : 1);
      int best = 1;
      long long best_cost = -1;
      for (int rpg = 1; rpg <= 12; ++rpg) {
        const long long groups = ((nx + rpg - 1) / rpg) * (long long)pl->gf.batch;
        const long long waves = (groups + slots - 1) / slots;
        const long long cost = waves * (rpg + 1 + (inverse ? 1 : 0));
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = rpg; }
      }
      return best;
    };
    const char* e = getenv("JAXIB_ROWS_PER_GROUP");
    pl->fast_rows_per_group = e ? atoi(e) : pick(false);
    pl->fast_rows_per_group_inv = e ? atoi(e) : pick(true);
    if (pl->fast_rows_per_group < 1) pl->fast_rows_per_group = 1;
    if (pl->fast_rows_per_group_inv < 1) pl->fast_rows_per_group_inv = 1;
    const char* ef = getenv("JAXIB_RPG_FWD");
    const char* ei = getenv("JAXIB_RPG_INV");
    if (ef) pl->fast_rows_per_group = atoi(ef);
    if (ei) pl->fast_rows_per_group_inv = atoi(ei);
  }
  const size_t col_bytes = (size_t)(padi(nx) + 1) * sizeof(float2);
  int cpc = 4;
  while (cpc > 1 && cpc * col_bytes > (size_t)max_smem) cpc /= 2;
  if (cpc * col_bytes > (size_t)max_smem) {
    set_error("nx=%d: one column (%zu B) exceeds shared memory", nx, col_bytes);
    jaxib_plan_destroy(pl);
    return JAXIB_ERR_UNSUPPORTED;
  }
  pl->cols_per_cta = cpc;
  pl->col_smem = cpc * col_bytes;
  pl->col_threads = 256;
  // The opt-in is a property of the FUNCTION, not of the plan: a per-plan value would be overwritten by the
  // next plan (a later, smaller plan would make the launches of a larger one fail).  Every generic kernel
  // is opted in to the device maximum once; the per-launch size still comes from the plan.
  if ((size_t)max_smem < pl->col_smem || (size_t)max_smem < 2 * pl->row_smem + 2 * (size_t)ny * sizeof(float)) {
    set_error("grid %dx%d: a row / column group does not fit the %d bytes of shared memory", nx, ny, max_smem);
    jaxib_plan_destroy(pl);
    return JAXIB_ERR_UNSUPPORTED;
  }
  cudaFuncSetAttribute(cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  cudaFuncSetAttribute(cols_dct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  cudaFuncSetAttribute(rows_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  cudaFuncSetAttribute(rows_inv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  cudaFuncSetAttribute(rows_inv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  cudaFuncSetAttribute(rows_fwd_dct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  cudaFuncSetAttribute(rows_inv_dct_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  cudaFuncSetAttribute(rows_inv_dct_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  cudaGetLastError();
  *out = pl;
  return JAXIB_OK;
}

extern "C" int jaxib_plan_destroy(jaxib_plan* pl) {
  if (!pl) return JAXIB_OK;
  cudaFree(pl->tw_row); cudaFree(pl->tw_col); cudaFree(pl->row_pos); cudaFree(pl->col_pos);
  cudaFree(pl->lam_x_pos); cudaFree(pl->lam_y); cudaFree(pl->lam_x_pos_f); cudaFree(pl->lam_y_f);
  cudaFree(pl->tw_dct); cudaFree(pl->tw_team_col); cudaFree(pl->tw_team_row);
  cudaFree(pl->tw_dct_col); cudaFree(pl->lam_x_nat);
  cudaFree(pl->cs_tab); cudaFree(pl->sw_rho); cudaFree(pl->sw_lg); cudaFree(pl->sw_gain); cudaFree(pl->sw_rho_d);
  delete pl;
  return JAXIB_OK;
}
