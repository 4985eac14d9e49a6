// K1: explicit terms and predictor on the staggered grid.
//
//   k   = convect(v) + nu * laplacian(v) + force/rho [- perm * v / rho]   (equations.py:42-83)
//   u*  = u + dt * k - forward_difference(p)                                (time_stepping.py:191-208)
//
// Advection: advect_upwind (advection.py:120-124), advect_van_leer (:206-280),
// advect_van_leer_using_limiters (:325-333 = interpolation.py:159-304).  Neighbour access follows
// ConstantBoundaryConditions.shift (boundaries.py:95-273): periodic wrap in x; in y either wrap or
// the Dirichlet ghost rules of a channel (cell-centred: 2*bc - mirror; edge-aligned: wall value
// when stepping onto the wall, bc - reflect beyond it).  Derived quantities (fluxes, the van Leer
// correction) obey their OWN ghost rule, inherited from the BC object the reference attaches to them.
//
// One CTA computes a TI x TJ tile from a shared-memory copy with a 2-cell halo; each HBM word of
// (u, v, p) is read once per tile (+ halo, served by L2) and each output word written once.
#include "common.cuh"

namespace jaxib {
namespace {

constexpr int TI = 16, TJ = 64, H = 2;
constexpr int PU = TJ + 2 * H + 1;  // padded pitch of the u/v tiles
constexpr int PP = TJ + 1 + 1;      // pitch of the p tile

__device__ __forceinline__ int wrap(int i, int n) {
  i %= n;
  return i < 0 ? i + n : i;
}

// Far field (Dirichlet on the x walls too, boundaries.py:666-693): the x-direction ghost rules are the y rules
// with the roles of u and v exchanged -- u is EDGE-aligned in x (offset 1: its last index lies on the upper
// wall), v is cell-centred in x.  `sgn`/`add` describe the affine map ghost = add + sgn * interior(i').
struct XGhost {
  int i;
  float sgn, add;
};
__device__ __forceinline__ XGhost xghost_edge(int i, int nx, float lo, float hi) {  // boundaries.py:205-246
  if (i == -1) return {0, 0.0f, lo};                       // the wall itself
  if (i < -1) return {-2 - i, -1.0f, lo};                   // bc - reflect past the wall
  if (i >= nx) return {2 * nx - 2 - i, -1.0f, hi};
  return {i, 1.0f, 0.0f};
}
__device__ __forceinline__ XGhost xghost_center(int i, int nx, float lo, float hi) {  // boundaries.py:179-188
  if (i < 0) return {-1 - i, -1.0f, 2.0f * lo};
  if (i >= nx) return {2 * nx - 1 - i, -1.0f, 2.0f * hi};
  return {i, 1.0f, 0.0f};
}

// u is cell-centred in y (offset .5): ghost = 2*bc - symmetric mirror (boundaries.py:179-188)
__device__ __forceinline__ float load_u_y(const float* __restrict__ u, const GridF& g, int i, int j) {
  if (j < 0) return 2.0f * g.uw_lo - u[(size_t)i * g.ny + (-1 - j)];
  if (j >= g.ny) return 2.0f * g.uw_hi - u[(size_t)i * g.ny + (2 * g.ny - 1 - j)];
  return u[(size_t)i * g.ny + j];
}
__device__ __forceinline__ float load_u(const float* __restrict__ u, const GridF& g, int i, int j) {
  if (g.bc == JAXIB_BC_FAR_FIELD) {
    const XGhost x = xghost_edge(i, g.nx, g.uwx_lo, g.uwx_hi);
    return x.sgn == 0.0f ? x.add : x.add + x.sgn * load_u_y(u, g, x.i, j);
  }
  i = wrap(i, g.nx);
  if (g.bc == JAXIB_BC_PERIODIC) return u[(size_t)i * g.ny + wrap(j, g.ny)];
  return load_u_y(u, g, i, j);
}

// v is edge-aligned in y (offset 1, last index on the upper wall): boundaries.py:189-246
__device__ __forceinline__ float load_v_y(const float* __restrict__ v, const GridF& g, int i, int j) {
  if (j == -1) return g.vw_lo;                                         // the wall itself
  if (j < -1) return g.vw_lo - v[(size_t)i * g.ny + (-2 - j)];         // bc - reflect past the wall
  if (j >= g.ny) return g.vw_hi - v[(size_t)i * g.ny + (2 * g.ny - 2 - j)];
  return v[(size_t)i * g.ny + j];
}
__device__ __forceinline__ float load_v(const float* __restrict__ v, const GridF& g, int i, int j) {
  if (g.bc == JAXIB_BC_FAR_FIELD) {
    const XGhost x = xghost_center(i, g.nx, g.vwx_lo, g.vwx_hi);
    return x.add + x.sgn * load_v_y(v, g, x.i, j);
  }
  i = wrap(i, g.nx);
  if (g.bc == JAXIB_BC_PERIODIC) return v[(size_t)i * g.ny + wrap(j, g.ny)];
  return load_v_y(v, g, i, j);
}

// p is cell-centred with homogeneous Neumann in y: ghost = edge (boundaries.py:247-269)
__device__ __forceinline__ float load_p(const float* __restrict__ p, const GridF& g, int i, int j) {
  if (g.bc == JAXIB_BC_FAR_FIELD) i = i < 0 ? 0 : (i >= g.nx ? g.nx - 1 : i);  // Neumann ghost = edge
  else i = wrap(i, g.nx);
  if (g.bc == JAXIB_BC_PERIODIC) return p[(size_t)i * g.ny + wrap(j, g.ny)];
  j = j < 0 ? 0 : (j >= g.ny ? g.ny - 1 : j);
  return p[(size_t)i * g.ny + j];
}

struct Tiles {
  const float* su;
  const float* sv;
  int i0, j0;  // global index of tile element [H][H]
  __device__ __forceinline__ float u(int i, int j) const { return su[(i - i0 + H) * PU + (j - j0 + H)]; }
  __device__ __forceinline__ float v(int i, int j) const { return sv[(i - i0 + H) * PU + (j - j0 + H)]; }
};

__device__ __forceinline__ float vl_corr(float cl, float cc, float cr) {
  // advection.py:258-264: 2 (c+ - c)(c - c-) / (c+ - c-) where the product is positive, else 0
  float prod = 2.0f * (cr - cc) * (cc - cl);
  return prod > 0.0f ? prod / (cr - cl) : 0.0f;
}

__device__ __forceinline__ float tvd_phi(float r) {
  // interpolation.py:225-232: van Leer limiter 2r / (1 + r) for r > 0 (safe_div)
  if (!(r > 0.0f)) return 0.0f;
  float den = 1.0f + r;
  return (2.0f * r) / (den != 0.0f ? den : 1.0f);
}

__device__ __forceinline__ float safe_div(float x, float y) { return x / (y != 0.0f ? y : 1.0f); }

// Face value of c for the flux through the face between `c0` (index k) and `cp` (index k+1), with the
// face velocity w.  cm = c[k-1], cpp = c[k+2].  corr_next is the van Leer correction at index k+1
// (already BC-resolved by the caller).
template <int SCHEME>
__device__ __forceinline__ float face_flux(float w, float cm, float c0, float cp, float cpp, float corr_next,
                                           float dt_over_h) {
  if (SCHEME == JAXIB_CONVECT_UPWIND) {
    return (w > 0.0f ? c0 : cp) * w;  // interpolation.py:148-152, advection.py:72
  } else if (SCHEME == JAXIB_CONVECT_VAN_LEER) {
    float up = w > 0.0f ? w * c0 : w * cp;
    float a = fabsf(w);
    float pre = 0.5f * (1.0f - dt_over_h * a) * a;
    float corr = w > 0.0f ? vl_corr(cm, c0, cp) : corr_next;
    return up + pre * corr;
  } else {  // Lax-Wendroff + TVD limiter
    float c_low = w > 0.0f ? c0 : cp;
    float courant = dt_over_h * w;
    float d = cp - c0;
    float c_high = w > 0.0f ? c0 + 0.5f * (1.0f - courant) * d : cp - 0.5f * (1.0f + courant) * d;
    float phi = w > 0.0f ? tvd_phi(safe_div(c0 - cm, d)) : tvd_phi(safe_div(cpp - cp, d));
    return (c_low - (c_low - c_high) * phi) * w;
  }
}

template <int SCHEME>
struct Terms {
  const GridF& g;
  const Tiles& t;
  const float* gu;  // the fields in global memory, for the wrapped upwind flux below
  const float* gv;
  __device__ Terms(const GridF& g_, const Tiles& t_, const float* gu_, const float* gv_) : g(g_), t(t_), gu(gu_), gv(gv_) {}

  // Plain upwind hands the interpolated variable back with PERIODIC boundary conditions (interpolation.py:152-156)
  // and the flux inherits them (advection.py:72-74), so the ghost flux below a wall is the flux of the LAST index
  // of that direction, wall or not.  It lies outside the tile: read its few inputs from global memory (one row /
  // one column of threads per step does).  The TVD and van Leer paths keep the scalar's own ghost rule.
  __device__ __forceinline__ float wrapped_u_fy(int i) const {
    const int j = g.ny - 1;
    const float w = 0.5f * load_v(gv, g, i, j) + 0.5f * load_v(gv, g, i + 1, j);
    return (w > 0.0f ? load_u(gu, g, i, j) : load_u(gu, g, i, j + 1)) * w;
  }
  __device__ __forceinline__ float wrapped_v_fy(int i) const {
    const int j = g.ny - 1;
    const float c0 = load_v(gv, g, i, j), cp = load_v(gv, g, i, j + 1);
    const float w = 0.5f * c0 + 0.5f * cp;
    return (w > 0.0f ? c0 : cp) * w;
  }
  __device__ __forceinline__ float wrapped_u_fx(int j) const {
    const int i = g.nx - 1;
    const float c0 = load_u(gu, g, i, j), cp = load_u(gu, g, i + 1, j);
    const float w = 0.5f * c0 + 0.5f * cp;
    return (w > 0.0f ? c0 : cp) * w;
  }
  __device__ __forceinline__ float wrapped_v_fx(int j) const {
    const int i = g.nx - 1;
    const float w = 0.5f * load_u(gu, g, i, j) + 0.5f * load_u(gu, g, i, j + 1);
    return (w > 0.0f ? load_v(gv, g, i, j) : load_v(gv, g, i + 1, j)) * w;
  }

  // ---- component u (offset (1, .5)) -------------------------------------------------------
  // The x-direction rules of the far field mirror the y-direction rules of the channel with u <-> v.
  __device__ __forceinline__ float u_corr_x(int i, int j) const {
    // correction at offset (1.5, .5), u's BC: cell-centred rule 2*bc - mirror (mirror of v_corr_y)
    if (g.bc == JAXIB_BC_FAR_FIELD && i >= g.nx) {
      int ii = 2 * g.nx - 1 - i;
      return 2.0f * g.uwx_hi - vl_corr(t.u(ii - 1, j), t.u(ii, j), t.u(ii + 1, j));
    }
    return vl_corr(t.u(i - 1, j), t.u(i, j), t.u(i + 1, j));
  }
  __device__ __forceinline__ float u_fx_raw(int i, int j) const {
    float w = 0.5f * t.u(i, j) + 0.5f * t.u(i + 1, j);
    float cn = SCHEME == JAXIB_CONVECT_VAN_LEER ? u_corr_x(i + 1, j) : 0.0f;
    return face_flux<SCHEME>(w, t.u(i - 1, j), t.u(i, j), t.u(i + 1, j), t.u(i + 2, j), cn, g.dt_over_dx);
  }
  __device__ __forceinline__ float u_fx(int i, int j) const {  // flux through the x-face i+1/2
    // far field: flux at offset (1.5, .5) with u's BC, cell-centred ghost 2*bc - mirror (mirror of v_fy)
    if (g.bc == JAXIB_BC_FAR_FIELD && i < 0)
      return SCHEME == JAXIB_CONVECT_UPWIND ? wrapped_u_fx(j) : 2.0f * g.uwx_lo - u_fx_raw(-1 - i, j);
    return u_fx_raw(i, j);
  }
  __device__ __forceinline__ float u_corr_y(int i, int j) const {
    // correction array carries v's BC at offset (1, 1): ghost = bc - reflect (boundaries.py:238-246)
    if (g.bc != JAXIB_BC_PERIODIC && j >= g.ny) {
      int jj = 2 * g.ny - 2 - j;
      return g.vw_hi - vl_corr(t.u(i, jj - 1), t.u(i, jj), t.u(i, jj + 1));
    }
    return vl_corr(t.u(i, j - 1), t.u(i, j), t.u(i, j + 1));
  }
  __device__ __forceinline__ float u_fy(int i, int j) const {  // flux through the y-face j+1/2
    if (g.bc != JAXIB_BC_PERIODIC && j < 0)                      // flux at offset (1,1) stepping onto the wall
      return SCHEME == JAXIB_CONVECT_UPWIND ? wrapped_u_fy(i) : g.uw_lo;
    float w = 0.5f * t.v(i, j) + 0.5f * t.v(i + 1, j);
    float cn = SCHEME == JAXIB_CONVECT_VAN_LEER ? u_corr_y(i, j + 1) : 0.0f;
    return face_flux<SCHEME>(w, t.u(i, j - 1), t.u(i, j), t.u(i, j + 1), t.u(i, j + 2), cn, g.dt_over_dy);
  }
  // ---- component v (offset (.5, 1)) -------------------------------------------------------
  __device__ __forceinline__ float v_corr_x(int i, int j) const {
    // correction array carries u's BC at offset (1, 1): ghost = bc - reflect (mirror of u_corr_y)
    if (g.bc == JAXIB_BC_FAR_FIELD && i >= g.nx) {
      int ii = 2 * g.nx - 2 - i;
      return g.uwx_hi - vl_corr(t.v(ii - 1, j), t.v(ii, j), t.v(ii + 1, j));
    }
    return vl_corr(t.v(i - 1, j), t.v(i, j), t.v(i + 1, j));
  }
  __device__ __forceinline__ float v_fx(int i, int j) const {
    if (g.bc == JAXIB_BC_FAR_FIELD && i < 0)                   // flux at offset (1, 1) stepping onto the wall (mirror of u_fy)
      return SCHEME == JAXIB_CONVECT_UPWIND ? wrapped_v_fx(j) : g.vwx_lo;
    float w = 0.5f * t.u(i, j) + 0.5f * t.u(i, j + 1);
    float cn = SCHEME == JAXIB_CONVECT_VAN_LEER ? v_corr_x(i + 1, j) : 0.0f;
    return face_flux<SCHEME>(w, t.v(i - 1, j), t.v(i, j), t.v(i + 1, j), t.v(i + 2, j), cn, g.dt_over_dx);
  }
  __device__ __forceinline__ float v_corr_y(int i, int j) const {
    // correction at offset (.5, 1.5), v's BC: cell-centred rule 2*bc - mirror (boundaries.py:179-188)
    if (g.bc != JAXIB_BC_PERIODIC && j >= g.ny) {
      int jj = 2 * g.ny - 1 - j;
      return 2.0f * g.vw_hi - vl_corr(t.v(i, jj - 1), t.v(i, jj), t.v(i, jj + 1));
    }
    return vl_corr(t.v(i, j - 1), t.v(i, j), t.v(i, j + 1));
  }
  __device__ __forceinline__ float v_fy_raw(int i, int j) const {
    float w = 0.5f * t.v(i, j) + 0.5f * t.v(i, j + 1);
    float cn = SCHEME == JAXIB_CONVECT_VAN_LEER ? v_corr_y(i, j + 1) : 0.0f;
    return face_flux<SCHEME>(w, t.v(i, j - 1), t.v(i, j), t.v(i, j + 1), t.v(i, j + 2), cn, g.dt_over_dy);
  }
  __device__ __forceinline__ float v_fy(int i, int j) const {
    // flux at offset (.5, 1.5) with v's BC: cell-centred ghost 2*bc - mirror
    if (g.bc != JAXIB_BC_PERIODIC && j < 0)
      return SCHEME == JAXIB_CONVECT_UPWIND ? wrapped_v_fy(i) : 2.0f * g.vw_lo - v_fy_raw(i, -1 - j);
    return v_fy_raw(i, j);
  }

  __device__ __forceinline__ float ku(int i, int j) const {
    float k = 0.0f;
    if (SCHEME != JAXIB_CONVECT_NONE)
      k = -((u_fx(i, j) - u_fx(i - 1, j)) * g.inv_dx + (u_fy(i, j) - u_fy(i, j - 1)) * g.inv_dy);
    if (g.has_nu) {
      float c = t.u(i, j);
      float lap = -2.0f * c * g.ssum + (t.u(i - 1, j) + t.u(i + 1, j)) * g.sx + (t.u(i, j - 1) + t.u(i, j + 1)) * g.sy;
      k += g.nu * lap;
    }
    return k;
  }
  __device__ __forceinline__ float kv(int i, int j) const {
    float k = 0.0f;
    if (SCHEME != JAXIB_CONVECT_NONE)
      k = -((v_fx(i, j) - v_fx(i - 1, j)) * g.inv_dx + (v_fy(i, j) - v_fy(i, j - 1)) * g.inv_dy);
    if (g.has_nu) {
      float c = t.v(i, j);
      float lap = -2.0f * c * g.ssum + (t.v(i - 1, j) + t.v(i + 1, j)) * g.sx + (t.v(i, j - 1) + t.v(i, j + 1)) * g.sy;
      k += g.nu * lap;
    }
    return k;
  }
};

template <int SCHEME, bool PREDICT>
__global__ void __launch_bounds__(256) explicit_kernel(GridF g, const float* __restrict__ u, const float* __restrict__ v,
                                                       const float* __restrict__ p, const float* __restrict__ perm,
                                                       float* __restrict__ out_u, float* __restrict__ out_v) {
  __shared__ float su[(TI + 2 * H) * PU];
  __shared__ float sv[(TI + 2 * H) * PU];
  __shared__ float sp[(TI + 1) * PP];
  const size_t plane = (size_t)g.nx * g.ny;
  const size_t boff = (size_t)blockIdx.z * plane;
  u += boff; v += boff; out_u += boff; out_v += boff;
  if (PREDICT && p) p += boff;
  if (perm) perm += boff;
  const int i0 = blockIdx.y * TI, j0 = blockIdx.x * TJ;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;

  for (int e = tid; e < (TI + 2 * H) * (TJ + 2 * H); e += 256) {
    int ti = e / (TJ + 2 * H), tj = e - ti * (TJ + 2 * H);
    int gi = i0 - H + ti, gj = j0 - H + tj;
    bool ok = gj >= -H && gj < g.ny + H;  // tiles overhanging the domain edge
    if (g.bc == JAXIB_BC_FAR_FIELD) ok = ok && gi >= -H && gi < g.nx + H;
    su[ti * PU + tj] = ok ? load_u(u, g, gi, gj) : 0.0f;
    sv[ti * PU + tj] = ok ? load_v(v, g, gi, gj) : 0.0f;
  }
  if (PREDICT && p) {
    for (int e = tid; e < (TI + 1) * (TJ + 1); e += 256) {
      int ti = e / (TJ + 1), tj = e - ti * (TJ + 1);
      sp[ti * PP + tj] = (j0 + tj <= g.ny) ? load_p(p, g, i0 + ti, j0 + tj) : 0.0f;
    }
  }
  __syncthreads();

  Tiles t{su, sv, i0, j0};
  Terms<SCHEME> terms(g, t, u, v);
  const int j = j0 + threadIdx.x;
  if (j >= g.ny) return;
#pragma unroll 1
  for (int r = threadIdx.y; r < TI; r += 4) {
    const int i = i0 + r;
    if (i >= g.nx) break;
    float ku = terms.ku(i, j) + g.fx;
    float kv = terms.kv(i, j) + g.fy;
    const size_t idx = (size_t)i * g.ny + j;
    if (perm) {  // penalty/util_funs.py:6-16: forcing = grad_p - perm * u, divided by density
      float k = perm[idx];
      ku -= k * t.u(i, j) * g.inv_rho;
      kv -= k * t.v(i, j) * g.inv_rho;
    }
    if (PREDICT) {
      float us = t.u(i, j) + g.dt * ku;
      float vs = t.v(i, j) + g.dt * kv;
      if (p) {
        const float* q = sp + r * PP + threadIdx.x;
        us -= (q[PP] - q[0]) * g.inv_dx;
        vs -= (q[1] - q[0]) * g.inv_dy;
      }
      out_u[idx] = us;
      out_v[idx] = vs;
    } else {
      out_u[idx] = ku;
      out_v[idx] = kv;
    }
  }
}

template <bool PREDICT>
int launch_scheme(cudaStream_t s, const GridF& g, const float* u, const float* v, const float* p, const float* perm,
                  float* out_u, float* out_v) {
  dim3 block(TJ, 4);
  dim3 grid((g.ny + TJ - 1) / TJ, (g.nx + TI - 1) / TI, g.batch);
  switch (g.convect) {
    case JAXIB_CONVECT_UPWIND:
      explicit_kernel<JAXIB_CONVECT_UPWIND, PREDICT><<<grid, block, 0, s>>>(g, u, v, p, perm, out_u, out_v);
      break;
    case JAXIB_CONVECT_VAN_LEER:
      explicit_kernel<JAXIB_CONVECT_VAN_LEER, PREDICT><<<grid, block, 0, s>>>(g, u, v, p, perm, out_u, out_v);
      break;
    case JAXIB_CONVECT_VAN_LEER_LIMITERS:
      explicit_kernel<JAXIB_CONVECT_VAN_LEER_LIMITERS, PREDICT><<<grid, block, 0, s>>>(g, u, v, p, perm, out_u, out_v);
      break;
    case JAXIB_CONVECT_NONE:
      explicit_kernel<JAXIB_CONVECT_NONE, PREDICT><<<grid, block, 0, s>>>(g, u, v, p, perm, out_u, out_v);
      break;
    default:
      set_error("unknown convect scheme %d", g.convect);
      return JAXIB_ERR_INVALID;
  }
  return check_launch("explicit_kernel");
}

}  // namespace

int launch_explicit(cudaStream_t s, const GridF& g, const float* u, const float* v, const float* p, const float* perm,
                    float* out_u, float* out_v, bool predictor) {
  if (g.nx < 4 || g.ny < 4) {
    set_error("grid %dx%d too small for the radius-2 stencil", g.nx, g.ny);
    return JAXIB_ERR_UNSUPPORTED;
  }
  int rc = JAXIB_OK;
  if (launch_explicit_ring(s, g, u, v, p, perm, out_u, out_v, predictor, &rc)) return rc;
  if (launch_explicit_fast(s, g, u, v, p, perm, out_u, out_v, predictor, &rc)) return rc;
  return predictor ? launch_scheme<true>(s, g, u, v, p, perm, out_u, out_v)
                   : launch_scheme<false>(s, g, u, v, p, perm, out_u, out_v);
}

}  // namespace jaxib
