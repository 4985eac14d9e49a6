// K1 fast path: periodic grid, first-order upwind advection (the Flapping / journal-bearing demos and
// the ib2048 bench workload).  Same arithmetic as stencil.cu (equations.py:42-83,
// time_stepping.py:191-208, advection.py:120-124) organised for bandwidth:
//
//   * each thread owns 4 consecutive y-cells (one float4 per field per row) and MARCHES along x over
//     RI rows, keeping rows i-1, i, i+1 of u and v (and i, i+1 of p) in registers;
//   * every HBM word of (u, v, p) is loaded once per CTA with 16-byte coalesced loads (+ 2 halo rows
//     per RI rows, served by L2); the one-cell y-halo comes from the neighbouring lane by shuffle
//     (warp-edge lanes read it directly);
//   * the x-face flux computed for row i is carried in registers as the (i-1)-face flux of row i+1,
//     so no flux is evaluated twice along x;
//   * the next row is prefetched into registers before the current row is computed.
#include "common.cuh"

namespace jaxib {
namespace {

constexpr int FT = 128;  // threads per CTA (each owns 4 columns)
constexpr unsigned FULLM = 0xffffffffu;

struct Row6 {  // columns j-1 .. j+4
  float a[6];
};
struct Row5 {  // columns j .. j+4
  float a[5];
};

__device__ __forceinline__ int wrap1(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

// loads columns j..j+3 (float4) of row `i` and completes the y-halo by shuffles / edge loads
__device__ __forceinline__ Row6 load_row6(const float* __restrict__ f, int i, int j, int ny, bool active, int lane) {
  const float* row = f + (size_t)i * ny;
  float4 c = active ? *reinterpret_cast<const float4*>(row + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  float left = __shfl_up_sync(FULLM, c.w, 1);
  float right = __shfl_down_sync(FULLM, c.x, 1);
  if (active) {
    if (lane == 0) left = row[j == 0 ? ny - 1 : j - 1];
    if (lane == 31 || j + 4 >= ny) right = row[j + 4 >= ny ? j + 4 - ny : j + 4];
  }
  Row6 r;
  r.a[0] = left; r.a[1] = c.x; r.a[2] = c.y; r.a[3] = c.z; r.a[4] = c.w; r.a[5] = right;
  return r;
}

__device__ __forceinline__ Row5 load_row5(const float* __restrict__ f, int i, int j, int ny, bool active, int lane) {
  const float* row = f + (size_t)i * ny;
  float4 c = active ? *reinterpret_cast<const float4*>(row + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  float right = __shfl_down_sync(FULLM, c.x, 1);
  if (active && (lane == 31 || j + 4 >= ny)) right = row[j + 4 >= ny ? j + 4 - ny : j + 4];
  Row5 r;
  r.a[0] = c.x; r.a[1] = c.y; r.a[2] = c.z; r.a[3] = c.w; r.a[4] = right;
  return r;
}

// Prefetch stage: the raw 16-byte column load plus the two warp-edge halo values.  The neighbour
// shuffles are deferred to `finish_row6/5` one iteration later -- doing them here would make the warp
// wait for the load it has just issued (ncu: 60 % long-scoreboard stalls on the SHFLs).
struct Raw6 {
  float4 c;
  float left, right;
};
__device__ __forceinline__ Raw6 fetch_row(const float* __restrict__ f, int i, int j, int ny, bool active, int lane) {
  const float* row = f + (size_t)i * ny;
  Raw6 r;
  r.c = make_float4(0.f, 0.f, 0.f, 0.f);
  r.left = r.right = 0.f;
  if (active) {
    r.c = *reinterpret_cast<const float4*>(row + j);
    if (lane == 0) r.left = row[j == 0 ? ny - 1 : j - 1];
    if (lane == 31 || j + 4 >= ny) r.right = row[j + 4 >= ny ? j + 4 - ny : j + 4];
  }
  return r;
}
__device__ __forceinline__ Row6 finish_row6(const Raw6& w, int j, int ny, bool active, int lane) {
  float left = __shfl_up_sync(FULLM, w.c.w, 1);
  float right = __shfl_down_sync(FULLM, w.c.x, 1);
  if (lane == 0) left = w.left;
  if (lane == 31 || j + 4 >= ny) right = w.right;
  Row6 r;
  r.a[0] = left; r.a[1] = w.c.x; r.a[2] = w.c.y; r.a[3] = w.c.z; r.a[4] = w.c.w; r.a[5] = right;
  return r;
}
__device__ __forceinline__ Row5 finish_row5(const Raw6& w, int j, int ny, bool active, int lane) {
  float right = __shfl_down_sync(FULLM, w.c.x, 1);
  if (lane == 31 || j + 4 >= ny) right = w.right;
  Row5 r;
  r.a[0] = w.c.x; r.a[1] = w.c.y; r.a[2] = w.c.z; r.a[3] = w.c.w; r.a[4] = right;
  return r;
}

__device__ __forceinline__ float upw(float w, float c0, float cp) { return (w > 0.0f ? c0 : cp) * w; }

template <bool PREDICT>
__global__ void __launch_bounds__(FT) explicit_upwind_periodic_kernel(GridF g, int rows_per_cta,
                                                                     const float* __restrict__ u,
                                                                     const float* __restrict__ v,
                                                                     const float* __restrict__ p,
                                                                     const float* __restrict__ perm,
                                                                     float* __restrict__ out_u,
                                                                     float* __restrict__ out_v) {
  pdl_trigger();
  pdl_wait();
  const size_t plane = (size_t)g.nx * g.ny;
  const size_t boff = (size_t)blockIdx.z * plane;
  u += boff; v += boff; out_u += boff; out_v += boff;
  if (p) p += boff;
  if (perm) perm += boff;
  const int lane = threadIdx.x & 31;
  const int j = (blockIdx.x * FT + threadIdx.x) * 4;
  const bool active = j < g.ny;
  const int i0 = blockIdx.y * rows_per_cta;
  const int i1 = min(i0 + rows_per_cta, g.nx);
  const int ny = g.ny, nx = g.nx;

  // window: rows i-1 (m), i (c), i+1 (n); `nn` is the prefetched row i+2
  Row6 um = load_row6(u, wrap1(i0 - 1, nx), j, ny, active, lane);
  Row6 vm = load_row6(v, wrap1(i0 - 1, nx), j, ny, active, lane);
  Row6 uc = load_row6(u, i0, j, ny, active, lane);
  Row6 vc = load_row6(v, i0, j, ny, active, lane);
  Row6 un = load_row6(u, wrap1(i0 + 1, nx), j, ny, active, lane);
  Row6 vn = load_row6(v, wrap1(i0 + 1, nx), j, ny, active, lane);
  Row5 pc, pn;
  if (PREDICT && p) {
    pc = load_row5(p, i0, j, ny, active, lane);
    pn = load_row5(p, wrap1(i0 + 1, nx), j, ny, active, lane);
  }
  // x-face fluxes through face (i0 - 1/2): between rows i0-1 and i0
  float fxu[4], fxv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    fxu[k] = upw(0.5f * um.a[k + 1] + 0.5f * uc.a[k + 1], um.a[k + 1], uc.a[k + 1]);
    fxv[k] = upw(0.5f * um.a[k + 1] + 0.5f * um.a[k + 2], vm.a[k + 1], vc.a[k + 1]);
  }

  for (int i = i0; i < i1; ++i) {
    // prefetch row i+2 (the next iteration's row i+1): loads only, no dependent instruction
    const int ip2 = wrap1(wrap1(i + 2, nx), nx);
    const Raw6 uraw = fetch_row(u, ip2, j, ny, active, lane);
    const Raw6 vraw = fetch_row(v, ip2, j, ny, active, lane);
    Raw6 praw;
    if (PREDICT && p) praw = fetch_row(p, ip2, j, ny, active, lane);

    // y-face fluxes of row i through faces k = 0..4 (face k sits between columns j-1+k and j+k)
    float fyu[5], fyv[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      fyu[k] = upw(0.5f * vc.a[k] + 0.5f * vn.a[k], uc.a[k], uc.a[k + 1]);
      fyv[k] = upw(0.5f * vc.a[k] + 0.5f * vc.a[k + 1], vc.a[k], vc.a[k + 1]);
    }
    float ou[4], ov[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = k + 1;
      // x-face fluxes through face (i + 1/2)
      const float fu = upw(0.5f * uc.a[c] + 0.5f * un.a[c], uc.a[c], un.a[c]);
      const float fv = upw(0.5f * uc.a[c] + 0.5f * uc.a[c + 1], vc.a[c], vn.a[c]);
      float ku = -((fu - fxu[k]) * g.inv_dx + (fyu[c] - fyu[c - 1]) * g.inv_dy);
      float kv = -((fv - fxv[k]) * g.inv_dx + (fyv[c] - fyv[c - 1]) * g.inv_dy);
      fxu[k] = fu;
      fxv[k] = fv;
      if (g.has_nu) {
        ku += g.nu * (-2.0f * uc.a[c] * g.ssum + (um.a[c] + un.a[c]) * g.sx + (uc.a[c - 1] + uc.a[c + 1]) * g.sy);
        kv += g.nu * (-2.0f * vc.a[c] * g.ssum + (vm.a[c] + vn.a[c]) * g.sx + (vc.a[c - 1] + vc.a[c + 1]) * g.sy);
      }
      ku += g.fx;
      kv += g.fy;
      if (PREDICT) {
        float us = uc.a[c] + g.dt * ku, vs = vc.a[c] + g.dt * kv;
        if (p) {
          us -= (pn.a[k] - pc.a[k]) * g.inv_dx;
          vs -= (pc.a[k + 1] - pc.a[k]) * g.inv_dy;
        }
        ou[k] = us;
        ov[k] = vs;
      } else {
        ou[k] = ku;
        ov[k] = kv;
      }
    }
    if (active) {
      const size_t idx = (size_t)i * ny + j;
      if (perm) {  // penalty/util_funs.py:6-16
        const float4 k4 = *reinterpret_cast<const float4*>(perm + idx);
        const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float s = PREDICT ? g.dt : 1.0f;
          ou[k] -= s * kk[k] * uc.a[k + 1] * g.inv_rho;
          ov[k] -= s * kk[k] * vc.a[k + 1] * g.inv_rho;
        }
      }
      *reinterpret_cast<float4*>(out_u + idx) = make_float4(ou[0], ou[1], ou[2], ou[3]);
      *reinterpret_cast<float4*>(out_v + idx) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    }
    um = uc; uc = un; un = finish_row6(uraw, j, ny, active, lane);
    vm = vc; vc = vn; vn = finish_row6(vraw, j, ny, active, lane);
    if (PREDICT && p) { pc = pn; pn = finish_row5(praw, j, ny, active, lane); }
  }
}

}  // namespace

// Returns true when the fast path handled the launch (rc holds the status).
bool launch_explicit_fast(cudaStream_t s, const GridF& g, const float* u, const float* v, const float* p,
                          const float* perm, float* out_u, float* out_v, bool predictor, int* rc) {
  if (g.bc != JAXIB_BC_PERIODIC || g.convect != JAXIB_CONVECT_UPWIND || (g.ny & 3) || g.nx < 4) return false;
  const uintptr_t align = reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(v) |
                          reinterpret_cast<uintptr_t>(out_u) | reinterpret_cast<uintptr_t>(out_v) |
                          reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(perm);
  if (align & 15) return false;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int strips = (g.ny / 4 + FT - 1) / FT;
  // enough CTAs for ~8 per SM, at least 4 rows each (halo-row overhead 2/RI)
  long long want = 8LL * sms;
  int rpc = (int)((long long)g.nx * strips * g.batch / want);
  rpc = rpc < 4 ? 4 : (rpc > 32 ? 32 : rpc);
  dim3 grid(strips, (g.nx + rpc - 1) / rpc, g.batch);
  if (predictor)
    launch_pdl(1, explicit_upwind_periodic_kernel<true>, grid, dim3(FT), 0, s, g, rpc, u, v, p, perm, out_u, out_v);
  else
    launch_pdl(1, explicit_upwind_periodic_kernel<false>, grid, dim3(FT), 0, s, g, rpc, u, v, (const float*)nullptr, perm, out_u, out_v);
  *rc = check_launch("explicit_upwind_periodic_kernel");
  return true;
}

}  // namespace jaxib
