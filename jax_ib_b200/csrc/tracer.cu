// Passive tracers and the Brinkman-penalisation mask.
//
//   bilinear sample   MD/CFD_MD_coupling.py:4-19  map_coordinates(order=1, mode='constant'):
//                     index = x/dx - offset, taps outside the array contribute 0 (no wrap)
//   tracer step       MD/simulate.py:81-97  dR = F dt/(m gamma) + sqrt(2 kT dt/(m gamma)) xi,
//                     F = pair_force(R) + [u, v](R); positions of immobile points are kept
//   pair force        MD/interaction_potential.py:6-22 through jax_md smap.pair + quantity.force:
//                     species-indexed (h, r0) harmonic wall / Morse tail, minimum-image displacement
//   penalty mask      penalty/util_funs.py:26-123: polar angle about the body centre, piecewise-
//                     linear R(theta), kappa = K/2 (1 + tanh((R^2 - r^2)/width)), summed over bodies
#include "common.cuh"

namespace jaxib {
namespace {

__device__ __forceinline__ float tap(const float* __restrict__ f, int nx, int ny, int i, int j) {
  return (i >= 0 && i < nx && j >= 0 && j < ny) ? __ldg(f + (size_t)i * ny + j) : 0.0f;
}

__device__ float bilinear(const float* __restrict__ f, const GridF& g, float off_x, float off_y, float x, float y) {
  const float ci = __fsub_rn(__fdiv_rn(x, g.dx), off_x);
  const float cj = __fsub_rn(__fdiv_rn(y, g.dy), off_y);
  const float fi0 = floorf(ci), fj0 = floorf(cj);
  const float wi = ci - fi0, wj = cj - fj0;
  const int i0 = (int)fi0, j0 = (int)fj0;
  float out = 0.0f;
  out = __fadd_rn(out, __fmul_rn(__fmul_rn(1.0f - wi, 1.0f - wj), tap(f, g.nx, g.ny, i0, j0)));
  out = __fadd_rn(out, __fmul_rn(__fmul_rn(1.0f - wi, wj), tap(f, g.nx, g.ny, i0, j0 + 1)));
  out = __fadd_rn(out, __fmul_rn(__fmul_rn(wi, 1.0f - wj), tap(f, g.nx, g.ny, i0 + 1, j0)));
  out = __fadd_rn(out, __fmul_rn(__fmul_rn(wi, wj), tap(f, g.nx, g.ny, i0 + 1, j0 + 1)));
  return out;
}

__global__ void sample_kernel(GridF g, float off_x, float off_y, const float* __restrict__ f,
                              const float* __restrict__ R, int n, float* __restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = bilinear(f, g, off_x, off_y, R[2 * k], R[2 * k + 1]);
}

__global__ void tracer_kernel(GridF g, const float* __restrict__ u, const float* __restrict__ v,
                              const float* __restrict__ pair_force, const float* __restrict__ xi,
                              const uint8_t* __restrict__ mobile, float dt, float nu_md, float noise, int n,
                              float* __restrict__ R) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  if (mobile && !mobile[k]) return;
  const float x = R[2 * k], y = R[2 * k + 1];
  float fx = bilinear(u, g, 1.0f, 0.5f, x, y);
  float fy = bilinear(v, g, 0.5f, 1.0f, x, y);
  if (pair_force) { fx = __fadd_rn(pair_force[2 * k], fx); fy = __fadd_rn(pair_force[2 * k + 1], fy); }
  float dxr = __fmul_rn(__fmul_rn(fx, dt), nu_md), dyr = __fmul_rn(__fmul_rn(fy, dt), nu_md);
  if (xi) { dxr = __fadd_rn(dxr, __fmul_rn(noise, xi[2 * k])); dyr = __fadd_rn(dyr, __fmul_rn(noise, xi[2 * k + 1])); }
  R[2 * k] = __fadd_rn(x, dxr);
  R[2 * k + 1] = __fadd_rn(y, dyr);
}

// One WARP per particle i (the first version ran one thread per particle: 1401 tracers = 44 warps on the whole GPU,
// 560 us per call at G2), tiles of 128 partners staged in shared memory; lane l takes partners l, l + 32, ... of
// every tile in index order and the 32 partial sums are combined by a fixed shuffle tree: deterministic.
constexpr int kPairWarps = 4;      // particles per CTA
constexpr int kPairSpecies = 8;    // species tables staged in shared memory up to this many species
__global__ void __launch_bounds__(32 * kPairWarps)
pair_force_kernel(const float* __restrict__ R, const int32_t* __restrict__ species, int n, int ns,
                  const float* __restrict__ h, const float* __restrict__ r0, float D0, float alpha,
                  float kk, float box_x, float box_y, float* __restrict__ F) {
  __shared__ float sx[128], sy[128];
  __shared__ int ssp[128];
  __shared__ float sh[kPairSpecies * kPairSpecies], sr[kPairSpecies * kPairSpecies];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * kPairWarps + warp;
  const bool act = i < n;
  const bool staged = ns <= kPairSpecies;
  if (staged)
    for (int e = threadIdx.x; e < ns * ns; e += blockDim.x) { sh[e] = h[e]; sr[e] = r0[e]; }
  const float xi = act ? R[2 * i] : 0.f, yi = act ? R[2 * i + 1] : 0.f;
  const int si = act ? species[i] : 0;
  float fx = 0.f, fy = 0.f;
  for (int base = 0; base < n; base += 128) {
    const int j = base + threadIdx.x;
    if (j < n) { sx[threadIdx.x] = R[2 * j]; sy[threadIdx.x] = R[2 * j + 1]; ssp[threadIdx.x] = species[j]; }
    __syncthreads();
    const int m = min(128, n - base);
    if (act) {
      for (int q = lane; q < m; q += 32) {
        if (base + q == i) continue;
        float dx = xi - sx[q], dy = yi - sy[q];
        dx -= box_x * rintf(dx / box_x);  // space.periodic displacement: minimum image
        dy -= box_y * rintf(dy / box_y);
        const float r = sqrtf(dx * dx + dy * dy);
        if (!(r > 0.f)) continue;
        const int e = si * ns + ssp[q];
        const float hh = staged ? sh[e] : h[e], rr = staged ? sr[e] : r0[e];
        float dU;
        if (r < rr) dU = 2.0f * hh * kk * (r - rr);
        else dU = D0 * (-2.0f * alpha * expf(-2.0f * alpha * (r - rr)) + 2.0f * alpha * expf(-alpha * (r - rr)));
        const float s = -dU / r;
        fx += s * dx;
        fy += s * dy;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    fx += __shfl_xor_sync(0xffffffffu, fx, o);
    fy += __shfl_xor_sync(0xffffffffu, fy, o);
  }
  if (act && lane == 0) { F[2 * i] = fx; F[2 * i + 1] = fy; }
}

__global__ void perm_kernel(GridF g, const float* __restrict__ centers, const float* __restrict__ Rtheta, int nb,
                            int nth, float K, float width, float* __restrict__ perm) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= g.ny || i >= g.nx) return;
  const float X = __fadd_rn(g.lower_x, __fmul_rn((float)i + 0.5f, g.dx));
  const float Y = __fadd_rn(g.lower_y, __fmul_rn((float)j + 0.5f, g.dy));
  const float pi = 3.14159265358979323846f;
  const float dth = (float)(2.0 * 3.14159265358979323846 / (nth - 1));
  float acc = 0.0f;
  for (int b = 0; b < nb; ++b) {
    const float xc = centers[2 * b], yc = centers[2 * b + 1];
    const float ddx = X - xc, ddy = Y - yc;
    const float at = atanf(ddy / ddx);
    // util_funs.py:38-42: three-branch polar angle in [0, 2 pi)
    float th = at * (ddx >= 0.f ? 1.f : 0.f) * (ddy > 0.f ? 1.f : 0.f);
    th += (at + pi) * (-ddx >= 0.f ? 1.f : 0.f);
    th += (at + 2.0f * pi) * (ddx > 0.f ? 1.f : 0.f) * (-ddy > 0.f ? 1.f : 0.f);
    int idx = (int)floorf(th / dth);
    idx = idx < 0 ? 0 : (idx > nth - 1 ? nth - 1 : idx);
    const int idn = idx + 1 < nth ? idx + 1 : 0;  // roll(-1)
    const float* Rt = Rtheta + (size_t)b * nth;
    const float th0 = (float)(idx * (2.0 * 3.14159265358979323846 / (nth - 1)));
    const float th1 = (float)(idn * (2.0 * 3.14159265358979323846 / (nth - 1)));
    const float slope = (Rt[idn] - Rt[idx]) / (th1 - th0);
    const float Rf = Rt[idx] + slope * (th - th0);
    const float G = Rf * Rf - (ddy * ddy + ddx * ddx);
    acc += K * 0.5f * (1.0f + tanhf(G / width));
  }
  perm[(size_t)i * g.ny + j] = acc;
}

}  // namespace
}  // namespace jaxib

using namespace jaxib;

extern "C" {

int jaxib_bilinear_sample(jaxib_stream_t stream, const jaxib_grid* grid, double off_x, double off_y,
                          const float* field, const float* R, int32_t n, float* out) {
  int rc = check_grid(grid);
  if (rc) return rc;
  if (n <= 0) return n == 0 ? JAXIB_OK : (set_error("bilinear_sample: n < 0"), JAXIB_ERR_INVALID);
  if (!field || !R || !out) { set_error("bilinear_sample: null device pointer"); return JAXIB_ERR_INVALID; }
  sample_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(make_gridf(*grid), (float)off_x, (float)off_y,
                                                                   field, R, n, out);
  return check_launch("sample_kernel");
}

int jaxib_tracer_step(jaxib_stream_t stream, const jaxib_grid* grid, const float* u, const float* v,
                      const float* pair_force, const float* xi, const uint8_t* is_mobile, double dt_md,
                      double mass_gamma, double kT, int32_t n, float* R) {
  int rc = check_grid(grid);
  if (rc) return rc;
  if (n <= 0) return n == 0 ? JAXIB_OK : (set_error("tracer_step: n < 0"), JAXIB_ERR_INVALID);
  if (!u || !v || !R) { set_error("tracer_step: null device pointer"); return JAXIB_ERR_INVALID; }
  if (!(mass_gamma > 0.0)) { set_error("tracer_step: mass*gamma must be positive"); return JAXIB_ERR_INVALID; }
  if (kT != 0.0 && !xi) { set_error("tracer_step: kT > 0 needs the normal deviates xi"); return JAXIB_ERR_INVALID; }
  const float dt = (float)dt_md;
  const float nu_md = 1.0f / (float)mass_gamma;
  const float noise = sqrtf(2.0f * (float)kT * dt * nu_md);
  tracer_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(make_gridf(*grid), u, v, pair_force,
                                                                   kT != 0.0 ? xi : nullptr, is_mobile, dt, nu_md,
                                                                   noise, n, R);
  return check_launch("tracer_kernel");
}

int jaxib_pair_force(jaxib_stream_t stream, const float* R, const int32_t* species, int32_t n, int32_t n_species,
                     const float* h, const float* r0, double D0, double alpha, double k, double box_x, double box_y,
                     float* F) {
  if (n <= 0) return n == 0 ? JAXIB_OK : (set_error("pair_force: n < 0"), JAXIB_ERR_INVALID);
  if (!R || !species || !h || !r0 || !F || n_species <= 0) { set_error("pair_force: bad argument"); return JAXIB_ERR_INVALID; }
  pair_force_kernel<<<(n + kPairWarps - 1) / kPairWarps, 32 * kPairWarps, 0, (cudaStream_t)stream>>>(R, species, n, n_species, h, r0, (float)D0,
                                                                       (float)alpha, (float)k, (float)box_x,
                                                                       (float)box_y, F);
  return check_launch("pair_force_kernel");
}

int jaxib_penalty_perm(jaxib_stream_t stream, const jaxib_grid* grid, const float* centers, const float* Rtheta,
                       int32_t n_bodies, int32_t ntheta, double K, double width, float* perm) {
  int rc = check_grid(grid);
  if (rc) return rc;
  if (!centers || !Rtheta || !perm || n_bodies <= 0 || ntheta < 2) { set_error("penalty_perm: bad argument"); return JAXIB_ERR_INVALID; }
  dim3 block(128), gridd((grid->ny + 127) / 128, grid->nx);
  perm_kernel<<<gridd, block, 0, (cudaStream_t)stream>>>(make_gridf(*grid), centers, Rtheta, n_bodies, ntheta,
                                                         (float)K, (float)width, perm);
  return check_launch("perm_kernel");
}

}  // extern "C"
