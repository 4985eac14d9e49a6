// K2/K3: immersed-boundary interpolation (E->L), direct forcing and spreading (L->E).
//
// Reference arithmetic (all float32):
//   delta(x, x0, w) = 1/(w sqrt(2 pi)) exp(-0.5 ((x - x0)/w)^2)              convolution_functions.py:5-7
//   U_k  = sum_ij field_ij * delta(xp_k, X_ij, dx) * delta(yp_k, Y_ij, dy) * dx * dy   :11-37
//   xp   = x0 cos(th) - y0 sin(th) + xc ; yp = x0 sin(th) + y0 cos(th) + yc   IBM_Force.py:33-34
//   UP_k = U0[Xi] + Omega * r_k, r = -(yp - yc) | (xp - xc) ; F_k = (UP_k - U_k)/dt      :39-50
//   dS_k = |x_{k+1} - x_k| (closed curve, roll(-1))                                       :59-63
//   f_ij = sum_k F_k * delta(sqrt((xp_k-X_ij)^2 + (yp_k-Y_ij)^2), 0, dx) * dS_k           :66-87
//   u**  = u* + dt * f                                                        time_stepping.py:223
//
// The reference evaluates every marker against the WHOLE grid; here each marker touches the
// 2R x 2R cells i in [i_k-R+1, i_k+R] around its owning cell i_k = floor(xp/dx - offset)
// (IBM_Force.py:35), R = 6 by default (truncation error 6e-9).  No periodic image is
// taken (the reference uses raw xp - X).
//
// Interpolation: one warp per marker; separable weights are computed once by 2R lanes and
// broadcast by shuffles; the window sum is a shuffle-tree reduction.
// Spreading: a deterministic GATHER -- one CTA per 16x16 cell tile of a body's bounding box, the
// body's markers are culled against the tile (order-preserving ballot compaction into shared
// memory) and every cell sums its contributions in marker order.  No float atomics anywhere.
// Marker coordinates and cell indices use explicitly rounded (non-contracted) float ops so that
// they match a NumPy float32 evaluation bit for bit.
#include <string.h>

#include "common.cuh"

namespace jaxib {
namespace {

constexpr int TS = 8;            // spread tile edge (cells)
constexpr int SPREAD_SLICES = 4;  // threads cooperating on one cell (measured: 2 slices 16.1 us, 4 slices 13.6 us at ib2048)
constexpr int SPREAD_THREADS = TS * TS * SPREAD_SLICES;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float mesh_coord(float lower, int i, float off, float h) {
  // grids.py:665: lower + (arange + offset) * step, each op rounded to float32
  return __fadd_rn(lower, __fmul_rn(__fadd_rn((float)i, off), h));
}
__device__ __forceinline__ int cell_index(float x, float lower, float h, float off) {
  // The cell whose mesh point lower + (i + offset) * h lies at or just below x: floor((x - lower)/h - offset).
  // IBM_Force.py:35 writes xp/dx - offset (its domains start at 0); x - 0 is exact, so the two agree bit for bit
  // there, and a domain with a non-zero origin keeps the window centred on the marker.
  return (int)floorf(__fsub_rn(__fdiv_rn(__fsub_rn(x, lower), h), off));
}
__device__ __forceinline__ float gauss(float x, float x0, float w, float pref) {
  float z = __fdiv_rn(__fsub_rn(x, x0), w);
  return __fmul_rn(pref, expf(__fmul_rn(-0.5f, __fmul_rn(z, z))));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

struct Mesh {
  float lower_x, lower_y, dx, dy, off_x, off_y, pref_x, pref_y;
  int nx, ny, R;
  int row0;  // global row of local row 0: indices are LOCAL, coordinates use local + row0
};

__device__ __forceinline__ Mesh make_mesh(const GridF& g, float off_x, float off_y) {
  Mesh m;
  m.lower_x = g.lower_x; m.lower_y = g.lower_y; m.dx = g.dx; m.dy = g.dy;
  m.off_x = off_x; m.off_y = off_y;
  const float s2pi = sqrtf(6.283185307179586f);
  m.pref_x = __fdiv_rn(1.0f, __fmul_rn(g.dx, s2pi));
  m.pref_y = __fdiv_rn(1.0f, __fmul_rn(g.dy, s2pi));
  m.nx = g.nx; m.ny = g.ny; m.R = g.R; m.row0 = g.row0;
  return m;
}

// Window sum for one marker, executed by a full warp.  Returns the sum in every lane.
// RT > 0 fixes the window radius at compile time: the window loop is unrolled and its loads issue together.
template <int RT>
__device__ __forceinline__ float interp_marker(const Mesh& m, const float* __restrict__ field, float xp, float yp,
                                               int* ik_out, int* jk_out) {
  const int lane = threadIdx.x & 31;
  const int R = RT > 0 ? RT : m.R;
  const int W = 2 * R;
  const int ik = cell_index(xp, m.lower_x, m.dx, m.off_x) - m.row0, jk = cell_index(yp, m.lower_y, m.dy, m.off_y);
  if (ik_out) { *ik_out = ik + m.row0; *jk_out = jk; }
  const int ia = ik - R + 1, ja = jk - R + 1;
  float gx = 0.0f, gy = 0.0f;
  if (lane < W) {
    gx = gauss(xp, mesh_coord(m.lower_x, ia + lane + m.row0, m.off_x, m.dx), m.dx, m.pref_x);
    gy = gauss(yp, mesh_coord(m.lower_y, ja + lane, m.off_y, m.dy), m.dy, m.pref_y);
  }
  float acc = 0.0f;
  const int cells = W * W;
#pragma unroll
  for (int base = 0; base < (RT > 0 ? 4 * RT * RT : cells); base += 32) {
    int e = base + lane;
    int wi = e / W, wj = e - wi * W;
    bool in = e < cells;
    float wx = __shfl_sync(FULL, gx, in ? wi : 0);
    float wy = __shfl_sync(FULL, gy, in ? wj : 0);
    int i = ia + wi, j = ja + wj;
    if (in && i >= 0 && i < m.nx && j >= 0 && j < m.ny) {
      float f = __ldg(field + (size_t)i * m.ny + j);
      // convolution_functions.py:19: ((((f * dx_delta) * dy_delta) * dx) * dy)
      acc += __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(f, wx), wy), m.dx), m.dy);
    }
  }
  return warp_sum(acc);
}

// ---- generic E->L interpolation (the `surface_fn` plug-in) -----------------------------------
__global__ void interp_points_kernel(GridF g, float off_x, float off_y, const float* __restrict__ field,
                                     const float* __restrict__ xp, const float* __restrict__ yp, int n,
                                     float* __restrict__ out, int32_t* __restrict__ ci, int32_t* __restrict__ cj) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= n) return;
  Mesh m = make_mesh(g, off_x, off_y);
  int ik, jk;
  float s = interp_marker<0>(m, field, xp[k], yp[k], &ik, &jk);
  if ((threadIdx.x & 31) == 0) {
    out[k] = s;
    if (ci) { ci[k] = ik; cj[k] = jk; }
  }
}

// ---- fused marker kinematics + interpolation + direct forcing --------------------------------
// scratch layout: [5][batch][n_markers] = xp, yp, dS, Fx, Fy
template <int RT>
__global__ void interp_force_kernel(GridF g, int n_bodies, int n_markers, const int32_t* __restrict__ body_offset,
                                    const float* __restrict__ x0, const float* __restrict__ y0,
                                    const float4* __restrict__ pack, const float* __restrict__ body_state,
                                    const float* __restrict__ u_star, const float* __restrict__ v_star,
                                    float* __restrict__ markers, int* __restrict__ tile_flags, int half_extent,
                                    int tiles_per_axis, int marker_first, int marker_count) {
  // one warp per (marker, velocity component): halves the dependent chain of the latency-bound kernel.
  // Dependent memory round trips: [x0, y0, pack] -> [body state] -> [window of u*] -> store.
  pdl_trigger();
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lw = wid >> 1, comp = wid & 1;
  // marker_count > 0: only the markers [marker_first, marker_first + marker_count) (slab decomposition: the bodies
  // near this slab, a host hint)
  const int mcount = marker_count > 0 ? marker_count : n_markers;
  if (lw >= g.batch * mcount) return;
  const int s = lw / mcount, k = marker_first + (lw - s * mcount);
  const int gw = s * n_markers + k;  // index into the marker arrays
  const int total = g.batch * n_markers;
  // body-frame data is constant over the run: fetched before the dependency wait (prologue under the
  // predictor's tail when launched programmatically)
  const float ax = __ldg(x0 + k), ay = __ldg(y0 + k);
  float bx, by;
  int b;
  if (pack) {
    const float4 pk = __ldg(pack + k);
    bx = pk.x; by = pk.y; b = __float_as_int(pk.z);
  } else {  // no pack: binary search of the owning body, then the next marker on its closed curve
    int lo = 0, hi = n_bodies;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (k >= __ldg(body_offset + mid)) lo = mid; else hi = mid;
    }
    b = lo;
    const int k0 = __ldg(body_offset + b), k1 = __ldg(body_offset + b + 1);
    const int kn = (k + 1 < k1) ? k + 1 : k0;  // roll(-1) within the body's closed curve
    bx = __ldg(x0 + kn); by = __ldg(y0 + kn);
  }
  // the kinematic table was uploaded before the block of steps started: not an output of the previous kernel
  const float4* st4 = reinterpret_cast<const float4*>(body_state + ((size_t)s * n_bodies + b) * JAXIB_BODY_STATE_STRIDE);
  const float4 sa = __ldg(st4), sb = __ldg(st4 + 1);
  pdl_wait();
  const float c = sa.x, sn = sa.y, xc = sa.z, yc = sa.w, U0x = sb.x, U0y = sb.y, om = sb.z;
  const float xp = __fadd_rn(__fsub_rn(__fmul_rn(ax, c), __fmul_rn(ay, sn)), xc);
  const float yp = __fadd_rn(__fadd_rn(__fmul_rn(ax, sn), __fmul_rn(ay, c)), yc);
  const float xq = __fadd_rn(__fsub_rn(__fmul_rn(bx, c), __fmul_rn(by, sn)), xc);
  const float yq = __fadd_rn(__fadd_rn(__fmul_rn(bx, sn), __fmul_rn(by, c)), yc);
  const size_t plane = (size_t)g.nx * g.ny;
  const Mesh m = comp == 0 ? make_mesh(g, 1.0f, 0.5f) : make_mesh(g, 0.5f, 1.0f);
  // slab decomposition: a slab only needs the forces of the markers near its rows (u-cell row in
  // [own_lo, own_hi)); the entries of the others are left untouched -- no cell of the slab sees them --
  // and their window sums are skipped (warp-uniform)
  const int iku = cell_index(xp, g.lower_x, g.dx, 1.0f);
  const bool wanted = g.own_hi <= g.own_lo || (iku >= g.own_lo && iku < g.own_hi);
  float U = 0.0f;
  if (wanted) U = interp_marker<RT>(m, (comp == 0 ? u_star : v_star) + s * plane, xp, yp, nullptr, nullptr);
  if (tile_flags != nullptr && wanted) {
    // mark the tiles of this body's bounding box (this component's mesh) that the marker's spreading window
    // touches: spread_bodies_kernel skips every other tile at once and clears the marks it consumes
    const int lane = threadIdx.x & 31;
    const int R = RT > 0 ? RT : g.R;
    const int ik = cell_index(xp, m.lower_x, m.dx, m.off_x) - m.row0, jk = cell_index(yp, m.lower_y, m.dy, m.off_y);
    const int ti0 = cell_index(xc, m.lower_x, m.dx, m.off_x) - m.row0 - half_extent, tj0 = cell_index(yc, m.lower_y, m.dy, m.off_y) - half_extent;
    const int ta = max((ik - R + 1 - ti0) >> 3, 0), tb = min((ik + R - ti0) >> 3, tiles_per_axis - 1);
    const int tc = max((jk - R + 1 - tj0) >> 3, 0), td = min((jk + R - tj0) >> 3, tiles_per_axis - 1);
    const int ny_t = td - tc + 1, nt = (tb - ta + 1) * ny_t;
    for (int e = lane; e < nt; e += 32) {
      const int ty = ta + e / ny_t, tx = tc + e % ny_t;
      tile_flags[(((size_t)s * n_bodies + b) * 2 + comp) * tiles_per_axis * tiles_per_axis + ty * tiles_per_axis + tx] = 1;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    const size_t stride = (size_t)total;
    float* mk = markers + gw;
    if (comp == 0) {
      const float rx = -(__fsub_rn(yp, yc));
      const float UPx = __fadd_rn(U0x, __fmul_rn(om, rx));
      const float dxl = __fsub_rn(xq, xp), dyl = __fsub_rn(yq, yp);
      mk[0] = xp;
      mk[stride] = yp;
      mk[2 * stride] = __fsqrt_rn(__fadd_rn(__fmul_rn(dxl, dxl), __fmul_rn(dyl, dyl)));
      if (wanted) mk[3 * stride] = __fdiv_rn(__fsub_rn(UPx, U), g.dt);
    } else {
      const float ry = __fsub_rn(xp, xc);
      const float UPy = __fadd_rn(U0y, __fmul_rn(om, ry));
      if (wanted) mk[4 * stride] = __fdiv_rn(__fsub_rn(UPy, U), g.dt);
    }
  }
}

// ---- spreading: tile gather ---------------------------------------------------------------------
constexpr int kCullBatches = 8;  // marker batches of SPREAD_THREADS the early cull handles (more: culled after the wait)
struct SpreadShared {
  float xp[SPREAD_THREADS], yp[SPREAD_THREADS], c[SPREAD_THREADS], part[SPREAD_THREADS];
  int ik[SPREAD_THREADS], jk[SPREAD_THREADS];
  int warp_count[SPREAD_THREADS / 32];
  int batch_end[kCullBatches + 1];  // early cull: list positions where each marker batch ends
};

// Culls markers [k0, k1) against the tile [ti0, ti0+TS) x [tj0, tj0+TS) (order-preserving ballot
// compaction) and accumulates cell (i, j).  A CTA is TS*TS cells x SPREAD_SLICES threads per cell:
// slice s sums list entries s, s + SLICES, ... in order and the partial sums are combined in a fixed
// order, so the result is deterministic (no atomics) while a body's few tiles still fill the SMs.
// idx != nullptr: the markers are idx[k0 .. k1) (ascending marker numbers, shared memory) instead of k0 .. k1 - 1.
__device__ float gather_markers(SpreadShared& sh, const Mesh& m, const float* __restrict__ mxp,
                                const float* __restrict__ myp, const float* __restrict__ mF,
                                const float* __restrict__ mdS, int k0, int k1, int ti0, int tj0, int i, int j,
                                bool valid, const int* idx = nullptr) {
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int slice = tid / (TS * TS);  // which quarter of the culled list this thread sums
  const float X = mesh_coord(m.lower_x, i + m.row0, m.off_x, m.dx);
  const float Y = mesh_coord(m.lower_y, j, m.off_y, m.dy);
  const float mhalf_inv_w2 = -0.5f / (m.dx * m.dx);
  float acc = 0.0f;
  for (int base = k0; base < k1; base += SPREAD_THREADS) {
    const int kk = base + tid;
    const int k = (idx != nullptr && kk < k1) ? idx[kk] : kk;
    bool keep = false;
    float xp = 0.f, yp = 0.f, c = 0.f;
    int ik = 0, jk = 0;
    if (kk < k1) {
      xp = mxp[k]; yp = myp[k];
      // IBM_Force.py:67: F * delta(r, 0, dx) * dS -- the marker's constant factor F * pref * dS
      c = mF[k] * m.pref_x * mdS[k];
      ik = cell_index(xp, m.lower_x, m.dx, m.off_x) - m.row0;
      jk = cell_index(yp, m.lower_y, m.dy, m.off_y);
      keep = (ik + m.R >= ti0) && (ik - m.R + 1 < ti0 + TS) && (jk + m.R >= tj0) && (jk - m.R + 1 < tj0 + TS);
    }
    const unsigned ballot = __ballot_sync(FULL, keep);
    if (lane == 0) sh.warp_count[warp] = __popc(ballot);
    __syncthreads();
    int offset = 0, totalc = 0;
#pragma unroll
    for (int w = 0; w < SPREAD_THREADS / 32; ++w) {
      int cw = sh.warp_count[w];
      if (w < warp) offset += cw;
      totalc += cw;
    }
    if (keep) {
      const int pos = offset + __popc(ballot & ((1u << lane) - 1u));
      sh.xp[pos] = xp; sh.yp[pos] = yp; sh.c[pos] = c; sh.ik[pos] = ik; sh.jk[pos] = jk;
    }
    __syncthreads();
    if (valid) {
      for (int q = slice; q < totalc; q += SPREAD_SLICES) {
        const int di = i - sh.ik[q], dj = j - sh.jk[q];
        if (di > -m.R && di <= m.R && dj > -m.R && dj <= m.R) {
          // delta(sqrt(d2), 0, dx) = pref * exp(-0.5 d2 / dx^2): width dx on both axes (IBM_Force.py:67)
          const float ddx = sh.xp[q] - X, ddy = sh.yp[q] - Y;
          acc = fmaf(sh.c[q], expf((ddx * ddx + ddy * ddy) * mhalf_inv_w2), acc);
        }
      }
    }
    __syncthreads();
  }
  // combine the slices in a fixed order
  sh.part[tid] = acc;
  __syncthreads();
  const int cell = tid & (TS * TS - 1);
  float tot = sh.part[cell];
#pragma unroll
  for (int q = 1; q < SPREAD_SLICES; ++q) tot += sh.part[cell + q * TS * TS];
  __syncthreads();
  return tot;  // meaningful in slice 0 (every slice computes the same value)
}

// Early cull (BEFORE the dependency wait): the marker positions follow from the body-frame coordinates and the
// kinematic table with the formulas of interp_force_kernel -- bit for bit what that kernel stores -- so a tile can
// build its list of markers while the force kernel is still running.  After the wait only the forces are missing:
// one round trip instead of three (marker coordinates, cull, forces).  The list keeps the batch structure of
// gather_markers (entries of marker batch n occupy [batch_end[n], batch_end[n+1])), so the later sum runs in exactly
// the same order and the result is bit-identical.  Returns the list length, or -1 if the list overflowed.
// sh.c holds dS * pref, sh.part (as int bits) the marker index.
__device__ int cull_markers_early(SpreadShared& sh, const Mesh& m, const float* __restrict__ x0,
                                  const float* __restrict__ y0, const float4* __restrict__ pack,
                                  const float* __restrict__ st, int k0, int k1, int ti0, int tj0) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float c = st[0], sn = st[1], xc = st[2], yc = st[3];
  int count = 0, nb = 0;
  bool overflow = false;
  if (tid == 0) sh.batch_end[0] = 0;
  for (int base = k0; base < k1; base += SPREAD_THREADS, ++nb) {
    const int k = base + tid;
    bool keep = false;
    float xp = 0.f, yp = 0.f, ds = 0.f;
    int ik = 0, jk = 0;
    if (k < k1) {
      const float ax = __ldg(x0 + k), ay = __ldg(y0 + k);
      const float4 pk = __ldg(pack + k);
      xp = __fadd_rn(__fsub_rn(__fmul_rn(ax, c), __fmul_rn(ay, sn)), xc);
      yp = __fadd_rn(__fadd_rn(__fmul_rn(ax, sn), __fmul_rn(ay, c)), yc);
      ik = cell_index(xp, m.lower_x, m.dx, m.off_x) - m.row0;
      jk = cell_index(yp, m.lower_y, m.dy, m.off_y);
      keep = (ik + m.R >= ti0) && (ik - m.R + 1 < ti0 + TS) && (jk + m.R >= tj0) && (jk - m.R + 1 < tj0 + TS);
      if (keep) {  // arc length to the next marker of the closed curve (interp_force_kernel, IBM_Force.py:59-63)
        const float xq = __fadd_rn(__fsub_rn(__fmul_rn(pk.x, c), __fmul_rn(pk.y, sn)), xc);
        const float yq = __fadd_rn(__fadd_rn(__fmul_rn(pk.x, sn), __fmul_rn(pk.y, c)), yc);
        const float dxl = __fsub_rn(xq, xp), dyl = __fsub_rn(yq, yp);
        ds = __fsqrt_rn(__fadd_rn(__fmul_rn(dxl, dxl), __fmul_rn(dyl, dyl)));
      }
    }
    const unsigned ballot = __ballot_sync(FULL, keep);
    if (lane == 0) sh.warp_count[warp] = __popc(ballot);
    __syncthreads();
    int offset = 0, totalc = 0;
#pragma unroll
    for (int w = 0; w < SPREAD_THREADS / 32; ++w) {
      const int cw = sh.warp_count[w];
      if (w < warp) offset += cw;
      totalc += cw;
    }
    if (count + totalc > SPREAD_THREADS) overflow = true;  // block-uniform
    if (keep && !overflow) {
      const int pos = count + offset + __popc(ballot & ((1u << lane) - 1u));
      sh.xp[pos] = xp; sh.yp[pos] = yp; sh.c[pos] = ds; sh.ik[pos] = ik; sh.jk[pos] = jk;
      sh.part[pos] = __int_as_float(k);
    }
    count += totalc;
    if (tid == 0) sh.batch_end[nb + 1] = count;
    __syncthreads();
  }
  return overflow ? -1 : count;
}

__device__ __forceinline__ void body_box(const float* st, const Mesh& m, int& ic, int& jc) {
  ic = cell_index(st[2], m.lower_x, m.dx, m.off_x) - m.row0;
  jc = cell_index(st[3], m.lower_y, m.dy, m.off_y);
}

// grid = (tiles_per_axis^2, n_bodies, batch * 2); block = SPREAD_THREADS; dynamic smem: (3 n_bodies + 1) ints
// Dependent memory round trips: [all body boxes + marker ranges -> smem, own cell of u* prefetched] -> [the
// body's markers] -> store (the first generation walked the bodies with one dependent global load each).
__global__ void __launch_bounds__(SPREAD_THREADS)
spread_bodies_kernel(GridF g, int n_bodies, int n_markers, const int32_t* __restrict__ body_offset,
                     const float* __restrict__ body_state, const float* __restrict__ markers, int half_extent,
                     int tiles_per_axis, float* __restrict__ u_star, float* __restrict__ v_star,
                     int* __restrict__ tile_flags, const float* __restrict__ x0, const float* __restrict__ y0,
                     const float4* __restrict__ pack, int early_cull, int body_first) {
  __shared__ SpreadShared sh;
  extern __shared__ int sbody[];  // ic[n_bodies] | jc[n_bodies] | offset[n_bodies + 1]
  int* const sic = sbody;
  int* const sjc = sbody + n_bodies;
  int* const soff = sbody + 2 * n_bodies;
  pdl_trigger();
  const int b = body_first + blockIdx.y;  // body_first > 0: the launch covers a range of bodies (slab hint)
  const int s = blockIdx.z >> 1, comp = blockIdx.z & 1;
  const Mesh m = comp == 0 ? make_mesh(g, 1.0f, 0.5f) : make_mesh(g, 0.5f, 1.0f);
  const float* states = body_state + (size_t)s * n_bodies * JAXIB_BODY_STATE_STRIDE;
  if (!early_cull && tile_flags != nullptr && n_bodies == 1) {
    // grids of many waves (ensembles): most CTAs start after the force kernel has finished -- the tile mark is the
    // cheapest test, an empty tile leaves after one load
    pdl_wait();
    if (tile_flags[(((size_t)s * n_bodies + b) * 2 + comp) * tiles_per_axis * tiles_per_axis + blockIdx.x] == 0) return;
  }
  // kinematic table and marker ranges are not outputs of the previous kernel: read before the dependency wait
  {  // a tile that lies off the grid leaves after one load (slab decomposition: most bodies belong to other slabs)
    int ic0, jc0;
    body_box(states + b * JAXIB_BODY_STATE_STRIDE, m, ic0, jc0);
    const int ty0 = blockIdx.x / tiles_per_axis, tx0 = blockIdx.x - ty0 * tiles_per_axis;
    const int ta = ic0 - half_extent + ty0 * TS, tb = jc0 - half_extent + tx0 * TS;
    if (ta >= g.nx || ta + TS <= 0 || tb >= g.ny || tb + TS <= 0) return;
  }
  for (int b2 = threadIdx.x; b2 <= n_bodies; b2 += SPREAD_THREADS) {
    if (b2 < n_bodies) {
      int ic2, jc2;
      body_box(states + b2 * JAXIB_BODY_STATE_STRIDE, m, ic2, jc2);
      sic[b2] = ic2;
      sjc[b2] = jc2;
    }
    soff[b2] = __ldg(body_offset + b2);
  }
  __syncthreads();
  const int ic = sic[b], jc = sjc[b];
  const int ty = blockIdx.x / tiles_per_axis, tx = blockIdx.x - ty * tiles_per_axis;
  const int ti0 = ic - half_extent + ty * TS, tj0 = jc - half_extent + tx * TS;
  if (ti0 >= g.nx || ti0 + TS <= 0 || tj0 >= g.ny || tj0 + TS <= 0) return;  // tile fully outside: no image (mark kept)
  // which OTHER bodies' boxes reach this tile?  One body per thread + a block-wide OR: with separated bodies
  // (the usual case) both answers are "none" and neither the ownership walk nor the multi-body gather runs
  bool lo_hit = false, hi_hit = false;
  for (int b2 = threadIdx.x; b2 < n_bodies; b2 += SPREAD_THREADS) {
    if (b2 == b) continue;
    const bool hit = !(sic[b2] + half_extent < ti0 || sic[b2] - half_extent >= ti0 + TS || sjc[b2] + half_extent < tj0 ||
                       sjc[b2] - half_extent >= tj0 + TS);
    lo_hit |= hit && b2 < b;
    hi_hit |= hit && b2 > b;
  }
  const bool any_lo = __syncthreads_or(lo_hit) != 0;
  const bool any_hi = __syncthreads_or(hi_hit) != 0;
  // early cull of this body's markers (see cull_markers_early): only when no later body reaches the tile
  int early = -1;
  if (early_cull && pack != nullptr && !any_hi && soff[b + 1] - soff[b] <= kCullBatches * SPREAD_THREADS) {
    early = cull_markers_early(sh, m, x0, y0, pack, states + b * JAXIB_BODY_STATE_STRIDE, soff[b], soff[b + 1], ti0, tj0);
    if (early == 0) return;  // no marker window touches the tile (the force kernel marks by the same test: no mark to clear)
  }
  // Occupied-tile marks (written by interp_force_kernel, the previous kernel, read below): a tile no marker of its
  // OWN body touches can only matter if another body's box overlaps it.
  const int cell = threadIdx.x & (TS * TS - 1);
  const int i = ti0 + cell / TS, j = tj0 + cell % TS;
  bool valid = i >= 0 && i < g.nx && j >= 0 && j < g.ny && i <= ic + half_extent && j <= jc + half_extent;
  // a cell covered by several bounding boxes belongs to the lowest-numbered body
  if (any_lo)
    for (int b2 = 0; b2 < b && valid; ++b2)
      if (abs(i - sic[b2]) <= half_extent && abs(j - sjc[b2]) <= half_extent) valid = false;
  const size_t total = (size_t)g.batch * n_markers;
  const float* mk = markers + (size_t)s * n_markers;
  const float* mF = mk + (comp == 0 ? 3 : 4) * total;
  float* const out = (comp == 0 ? u_star : v_star) + (size_t)s * g.nx * g.ny + (size_t)(valid ? i : 0) * g.ny + (valid ? j : 0);
  const bool writer = valid && threadIdx.x < TS * TS;
  int* const my_flag = tile_flags == nullptr ? nullptr :
      tile_flags + (((size_t)s * n_bodies + b) * 2 + comp) * tiles_per_axis * tiles_per_axis + blockIdx.x;
  pdl_wait();
  // one round trip after the wait: the tile's mark, the cell's current value and (early cull) the listed forces
  const int mark = my_flag != nullptr ? *my_flag : 1;
  const float cur = writer ? *out : 0.0f;
  float force = 0.0f;
  if (early > 0 && (int)threadIdx.x < early) force = mF[__float_as_int(sh.part[threadIdx.x])];
  if (my_flag != nullptr) {
    __syncthreads();  // every thread has read the mark before it is cleared
    if (mark == 0) {
      if (!any_hi) return;  // nothing of this body here (slab mode: its markers belong to another slab) and no later body reaches the tile
    } else if (threadIdx.x == 0) {
      *my_flag = 0;  // consumed: all marks are zero again after the step
    }
  }
  float f_total = 0.0f;
  bool touched = false;
  if (early > 0) {
    // the list is ready: fetch the forces, then sum batch by batch exactly as gather_markers does
    const int tid = threadIdx.x;
    if (tid < early) sh.c[tid] = force * m.pref_x * sh.c[tid];  // F * pref * dS (IBM_Force.py:67)
    __syncthreads();
    const int slice = tid / (TS * TS);
    const float X = mesh_coord(m.lower_x, i + m.row0, m.off_x, m.dx);
    const float Y = mesh_coord(m.lower_y, j, m.off_y, m.dy);
    const float mhalf_inv_w2 = -0.5f / (m.dx * m.dx);
    float acc = 0.0f;
    const int nb = (soff[b + 1] - soff[b] + SPREAD_THREADS - 1) / SPREAD_THREADS;
    if (valid) {
      for (int n = 0; n < nb; ++n) {
        const int q0 = sh.batch_end[n], q1 = sh.batch_end[n + 1];
        for (int q = q0 + slice; q < q1; q += SPREAD_SLICES) {
          const int di = i - sh.ik[q], dj = j - sh.jk[q];
          if (di > -m.R && di <= m.R && dj > -m.R && dj <= m.R) {
            const float ddx = sh.xp[q] - X, ddy = sh.yp[q] - Y;
            acc = fmaf(sh.c[q], expf((ddx * ddx + ddy * ddy) * mhalf_inv_w2), acc);
          }
        }
      }
    }
    __syncthreads();  // sh.part is reused for the partial sums
    sh.part[tid] = acc;
    __syncthreads();
    const int cell2 = tid & (TS * TS - 1);
    float tot = sh.part[cell2];
#pragma unroll
    for (int q = 1; q < SPREAD_SLICES; ++q) tot += sh.part[cell2 + q * TS * TS];
    f_total = __fadd_rn(f_total, tot);
    touched = true;
  } else {
    const int b_end = any_hi ? n_bodies : b + 1;
    for (int b2 = b; b2 < b_end; ++b2) {  // IBM_Force.py:99-105: force += per-body contribution
      const int ic2 = sic[b2], jc2 = sjc[b2];
      if (b2 != b && (ic2 + half_extent < ti0 || ic2 - half_extent >= ti0 + TS || jc2 + half_extent < tj0 ||
                      jc2 - half_extent >= tj0 + TS))
        continue;  // block-uniform
      float fb = gather_markers(sh, m, mk, mk + total, mF, mk + 2 * total, soff[b2], soff[b2 + 1], ti0, tj0, i, j, valid);
      f_total = __fadd_rn(f_total, fb);
      touched = true;
    }
  }
  if (writer && touched && f_total != 0.0f) *out = __fadd_rn(cur, __fmul_rn(g.dt, f_total));  // time_stepping.py:223
}

// ---- generic L->E spreading of a point cloud (plug-in granularity, batch = 1) -------------------
__global__ void bbox_kernel(GridF g, float off_x, float off_y, const float* __restrict__ xp,
                            const float* __restrict__ yp, int n, int* __restrict__ box) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int ik = cell_index(xp[k], g.lower_x, g.dx, off_x) - g.row0, jk = cell_index(yp[k], g.lower_y, g.dy, off_y);
  atomicMin(box + 0, ik); atomicMax(box + 1, ik);
  atomicMin(box + 2, jk); atomicMax(box + 3, jk);
}
__global__ void bbox_init_kernel(int* box) {
  box[0] = box[2] = 0x7fffffff;
  box[1] = box[3] = (int)0x80000000;
}

// grid = (ceil(ny/TS), ceil(nx/TS)); every tile culls all n points (early exit outside the bbox)
__global__ void __launch_bounds__(SPREAD_THREADS)
spread_points_kernel(GridF g, float off_x, float off_y, const float* __restrict__ F, const float* __restrict__ xp,
                     const float* __restrict__ yp, const float* __restrict__ dS, int n, float scale,
                     const int* __restrict__ box, float* __restrict__ out) {
  __shared__ SpreadShared sh;
  const int ti0 = blockIdx.y * TS, tj0 = blockIdx.x * TS;
  if (ti0 > box[1] + g.R || ti0 + TS <= box[0] - g.R || tj0 > box[3] + g.R || tj0 + TS <= box[2] - g.R) return;
  const Mesh m = make_mesh(g, off_x, off_y);
  const int cell = threadIdx.x & (TS * TS - 1);
  const int i = ti0 + cell / TS, j = tj0 + cell % TS;
  const bool valid = i < g.nx && j < g.ny;
  float f = gather_markers(sh, m, xp, yp, F, dS, 0, n, ti0, tj0, i, j, valid);
  if (valid && f != 0.0f && threadIdx.x < TS * TS) {
    float* o = out + (size_t)i * g.ny + j;
    *o = __fadd_rn(*o, __fmul_rn(scale, f));
  }
}

// ---- cell-binned spreading of a point cloud ---------------------------------------------------------------------
// Markers are counting-sorted into bins of BN x BN cells by their owning cell (IBM_Force.py:35); a tile then only
// looks at the markers of the few bins its spreading window can reach instead of culling all n points: the cost is
// O(markers x window), independent of how many tiles the cloud covers.  The fill order inside a bin depends on the
// atomics' arrival order, so every tile SORTS its candidate marker numbers before summing: the sum runs in marker
// order and the result is deterministic (and no float is ever added atomically).
constexpr int BN = 32;           // bin edge in cells
constexpr int kCandCap = 2048;   // candidate markers a tile can sort; more: the tile falls back to culling all n points

__device__ __forceinline__ int bin_of(int ik, int jk, int bnx, int bny) {
  const int bi = min(max(ik, 0) / BN, bnx - 1), bj = min(max(jk, 0) / BN, bny - 1);  // markers off the grid: nearest bin
  return bi * bny + bj;
}

__global__ void bin_count_kernel(GridF g, float off_x, float off_y, const float* __restrict__ xp,
                                 const float* __restrict__ yp, int n, int bnx, int bny, int* __restrict__ count) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int ik = cell_index(xp[k], g.lower_x, g.dx, off_x) - g.row0, jk = cell_index(yp[k], g.lower_y, g.dy, off_y);
  atomicAdd(count + bin_of(ik, jk, bnx, bny), 1);
}

// exclusive scan of the bin counts by ONE CTA; offset[nbins] = n; cursor = copy of offset for the fill
__global__ void __launch_bounds__(1024) bin_scan_kernel(int nbins, const int* __restrict__ count, int* __restrict__ offset,
                                                        int* __restrict__ cursor) {
  __shared__ int wsum[32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int per = (nbins + 1023) / 1024;
  const int lo = min(t * per, nbins), hi = min(lo + per, nbins);
  int s = 0;
  for (int b = lo; b < hi; ++b) s += count[b];
  int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  int acc = inc - s;
  for (int q = 0; q < w; ++q) acc += wsum[q];
  for (int b = lo; b < hi; ++b) {
    offset[b] = acc;
    cursor[b] = acc;
    acc += count[b];
  }
  if (lo < hi && hi == nbins) offset[nbins] = acc;  // the thread that owns the last bin
}

__global__ void bin_fill_kernel(GridF g, float off_x, float off_y, const float* __restrict__ xp,
                                const float* __restrict__ yp, int n, int bnx, int bny, int* __restrict__ cursor,
                                int* __restrict__ list) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int ik = cell_index(xp[k], g.lower_x, g.dx, off_x) - g.row0, jk = cell_index(yp[k], g.lower_y, g.dy, off_y);
  list[atomicAdd(cursor + bin_of(ik, jk, bnx, bny), 1)] = k;
}

// grid = (ceil(ny/TS), ceil(nx/TS))
__global__ void __launch_bounds__(SPREAD_THREADS)
spread_binned_kernel(GridF g, float off_x, float off_y, const float* __restrict__ F, const float* __restrict__ xp,
                     const float* __restrict__ yp, const float* __restrict__ dS, int n, float scale,
                     const int* __restrict__ box, const int* __restrict__ offset, const int* __restrict__ list, int bnx,
                     int bny, float* __restrict__ out) {
  __shared__ SpreadShared sh;
  __shared__ int cand[kCandCap];
  const int ti0 = blockIdx.y * TS, tj0 = blockIdx.x * TS;
  if (ti0 > box[1] + g.R || ti0 + TS <= box[0] - g.R || tj0 > box[3] + g.R || tj0 + TS <= box[2] - g.R) return;
  const Mesh m = make_mesh(g, off_x, off_y);
  const int tid = threadIdx.x;
  const int cell = tid & (TS * TS - 1);
  const int i = ti0 + cell / TS, j = tj0 + cell % TS;
  const bool valid = i < g.nx && j < g.ny;
  // bins whose cells can own a marker that reaches this tile: owning cell in [t0 - R, t0 + TS - 1 + R - 1]
  // (edge bins also hold the markers off the grid)
  const int bi0 = min(max(ti0 - g.R, 0) / BN, bnx - 1), bi1 = min(max(ti0 + TS + g.R - 2, 0) / BN, bnx - 1);
  const int bj0 = min(max(tj0 - g.R, 0) / BN, bny - 1), bj1 = min(max(tj0 + TS + g.R - 2, 0) / BN, bny - 1);
  int ncand = 0;
  for (int bi = bi0; bi <= bi1; ++bi)
    for (int bj = bj0; bj <= bj1; ++bj) {
      const int o0 = offset[bi * bny + bj], o1 = offset[bi * bny + bj + 1];
      for (int q = o0 + tid; q < o1; q += SPREAD_THREADS)
        if (ncand + q - o0 < kCandCap) cand[ncand + q - o0] = list[q];
      ncand += o1 - o0;  // block-uniform
    }
  float f;
  if (ncand > kCandCap) {
    f = gather_markers(sh, m, xp, yp, F, dS, 0, n, ti0, tj0, i, j, valid);  // crowded tile: cull all points
  } else {
    int np2 = 1;
    while (np2 < ncand) np2 <<= 1;
    __syncthreads();
    for (int q = ncand + tid; q < np2; q += SPREAD_THREADS) cand[q] = 0x7fffffff;
    __syncthreads();
    for (int kk = 2; kk <= np2; kk <<= 1)  // bitonic sort, ascending marker numbers
      for (int jj = kk >> 1; jj > 0; jj >>= 1) {
        for (int q = tid; q < np2; q += SPREAD_THREADS) {
          const int p = q ^ jj;
          if (p > q) {
            const int a = cand[q], b = cand[p];
            const bool up = (q & kk) == 0;
            if ((a > b) == up) { cand[q] = b; cand[p] = a; }
          }
        }
        __syncthreads();
      }
    f = gather_markers(sh, m, xp, yp, F, dS, 0, ncand, ti0, tj0, i, j, valid, cand);
  }
  if (valid && f != 0.0f && tid < TS * TS) {
    float* o = out + (size_t)i * g.ny + j;
    *o = __fadd_rn(*o, __fmul_rn(scale, f));
  }
}

}  // namespace

// bounding box of a body in cells (half edge) and 8 x 8 tiles per axis: shared by the two fused IB kernels
static void spread_geometry(const GridF& g, const jaxib_bodies& b, int* half_extent, int* tiles) {
  const float hmin = g.dx < g.dy ? g.dx : g.dy;
  *half_extent = (int)ceil(b.max_radius / hmin) + g.R + 2;
  *tiles = (2 * *half_extent + 1 + TS - 1) / TS;
}

int ib_half_extent(const GridF& g, const jaxib_bodies& b) {
  int he, tiles;
  spread_geometry(g, b, &he, &tiles);
  return he;
}

int launch_ib_interp(cudaStream_t s, const GridF& g, float off_x, float off_y, const float* field, const float* xp,
                     const float* yp, int n, float* out, int32_t* ci, int32_t* cj) {
  if (n <= 0) return JAXIB_OK;
  if (g.R < 1 || g.R > 16) { set_error("ib_window R=%d outside [1,16]", g.R); return JAXIB_ERR_UNSUPPORTED; }
  const int threads = 128;
  const long long total = (long long)n * 32;
  interp_points_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(g, off_x, off_y, field, xp, yp,
                                                                                       n, out, ci, cj);
  return check_launch("interp_points_kernel");
}

int launch_ib_marker_force(cudaStream_t s, const GridF& g, const jaxib_bodies& b, const float* body_state,
                           const float* u_star, const float* v_star, float* markers, const BodyRange* range) {
  if (b.n_markers <= 0 || b.n_bodies <= 0) return JAXIB_OK;
  if (g.R < 1 || g.R > 16) { set_error("ib_window R=%d outside [1,16]", g.R); return JAXIB_ERR_UNSUPPORTED; }
  const int threads = 128;
  const int m_first = range && range->marker_count > 0 ? range->marker_first : 0;
  const int m_count = range && range->marker_count > 0 ? range->marker_count : 0;
  const long long total = (long long)g.batch * (m_count > 0 ? m_count : b.n_markers) * 2 * 32;
  const unsigned nblk = (unsigned)((total + threads - 1) / threads);
  int half_extent, tiles;
  spread_geometry(g, b, &half_extent, &tiles);
  const long long need = (long long)g.batch * b.n_bodies * 2 * tiles * tiles;
  int* flags = (b.tile_flags && (long long)b.tile_flags_len >= need) ? b.tile_flags : nullptr;
  if (g.R == 6)
    launch_pdl(2, interp_force_kernel<6>, dim3(nblk), dim3(threads), 0, s, g, b.n_bodies, b.n_markers, b.body_offset, b.x0,
               b.y0, reinterpret_cast<const float4*>(b.marker_pack), body_state, u_star, v_star, markers, flags, half_extent,
               tiles, m_first, m_count);
  else
    launch_pdl(2, interp_force_kernel<0>, dim3(nblk), dim3(threads), 0, s, g, b.n_bodies, b.n_markers, b.body_offset, b.x0,
               b.y0, reinterpret_cast<const float4*>(b.marker_pack), body_state, u_star, v_star, markers, flags, half_extent,
               tiles, m_first, m_count);
  return check_launch("interp_force_kernel");
}

int launch_ib_spread_bodies(cudaStream_t s, const GridF& g, const jaxib_bodies& b, const float* body_state,
                            const float* markers, float* u_star, float* v_star, const BodyRange* range) {
  if (b.n_markers <= 0 || b.n_bodies <= 0) return JAXIB_OK;
  if (g.R < 1 || g.R > 16) { set_error("ib_window R=%d outside [1,16]", g.R); return JAXIB_ERR_UNSUPPORTED; }
  int half_extent, tiles;
  spread_geometry(g, b, &half_extent, &tiles);
  const int b_first = range && range->body_count > 0 ? range->body_first : 0;
  const int b_count = range && range->body_count > 0 ? range->body_count : b.n_bodies;
  dim3 grid(tiles * tiles, b_count, g.batch * 2), block(SPREAD_THREADS);
  const size_t smem = (size_t)(3 * b.n_bodies + 1) * sizeof(int);
  if (smem > 40 * 1024) { set_error("ib_spread_bodies: %d bodies exceed the shared-memory body table", b.n_bodies); return JAXIB_ERR_UNSUPPORTED; }
  const long long need = (long long)g.batch * b.n_bodies * 2 * tiles * tiles;
  int* flags = (b.tile_flags && (long long)b.tile_flags_len >= need) ? b.tile_flags : nullptr;
  // The early cull pays while the CTAs sit resident under the force kernel (a few hundred tiles).  A grid of many
  // waves (ensembles: 128 x 64 tiles x 2) starts most CTAs after the force kernel has finished: there the tile mark
  // is the cheaper test (measured, ens1024: 74 us with the mark first, 113 us with the cull first).
  const int early_cull = (long long)grid.x * grid.y * grid.z <= 4096 ? 1 : 0;
  launch_pdl(4, spread_bodies_kernel, grid, block, smem, s, g, b.n_bodies, b.n_markers, b.body_offset, body_state, markers,
             half_extent, tiles, u_star, v_star, flags, b.x0, b.y0, reinterpret_cast<const float4*>(b.marker_pack), early_cull, b_first);
  return check_launch("spread_bodies_kernel");
}

int launch_ib_force(cudaStream_t s, const GridF& g, const jaxib_bodies& b, const float* body_state, float* u_star,
                    float* v_star, float* markers, const StageHook* hook) {
  int rc = launch_ib_marker_force(s, g, b, body_state, u_star, v_star, markers);
  if (rc) return rc;
  if (hook) cudaEventRecord(hook->ev[0], s);
  rc = launch_ib_spread_bodies(s, g, b, body_state, markers, u_star, v_star);
  if (hook) cudaEventRecord(hook->ev[1], s);
  return rc;
}

size_t ib_spread_scratch_bytes(const GridF& g, int n) {
  const size_t nbins = (size_t)((g.nx + BN - 1) / BN) * ((g.ny + BN - 1) / BN);
  return 16 + (3 * nbins + 1 + (size_t)(n > 0 ? n : 0)) * sizeof(int);
}

int launch_ib_spread(cudaStream_t s, const GridF& g, float off_x, float off_y, const float* F, const float* xp,
                     const float* yp, const float* dS, int n, float scale, float* out, int* box_scratch,
                     size_t scratch_bytes) {
  if (n <= 0) return JAXIB_OK;
  if (g.R < 1 || g.R > 16) { set_error("ib_window R=%d outside [1,16]", g.R); return JAXIB_ERR_UNSUPPORTED; }
  bbox_init_kernel<<<1, 1, 0, s>>>(box_scratch);
  bbox_kernel<<<(n + 255) / 256, 256, 0, s>>>(g, off_x, off_y, xp, yp, n, box_scratch);
  dim3 grid((g.ny + TS - 1) / TS, (g.nx + TS - 1) / TS), block(SPREAD_THREADS);
  if (scratch_bytes >= ib_spread_scratch_bytes(g, n)) {
    // cell-binned path: counting sort of the markers by bin, tiles gather from the bins in reach
    const int bnx = (g.nx + BN - 1) / BN, bny = (g.ny + BN - 1) / BN, nbins = bnx * bny;
    int* count = box_scratch + 4;
    int* offset = count + nbins;
    int* cursor = offset + nbins + 1;
    int* list = cursor + nbins;
    cudaMemsetAsync(count, 0, (size_t)nbins * sizeof(int), s);
    bin_count_kernel<<<(n + 255) / 256, 256, 0, s>>>(g, off_x, off_y, xp, yp, n, bnx, bny, count);
    bin_scan_kernel<<<1, 1024, 0, s>>>(nbins, count, offset, cursor);
    bin_fill_kernel<<<(n + 255) / 256, 256, 0, s>>>(g, off_x, off_y, xp, yp, n, bnx, bny, cursor, list);
    spread_binned_kernel<<<grid, block, 0, s>>>(g, off_x, off_y, F, xp, yp, dS, n, scale, box_scratch, offset, list, bnx,
                                                bny, out);
    return check_launch("spread_binned_kernel");
  }
  // minimal scratch (16 bytes): every tile inside the bounding box culls all n points
  spread_points_kernel<<<grid, block, 0, s>>>(g, off_x, off_y, F, xp, yp, dS, n, scale, box_scratch, out);
  return check_launch("spread_points_kernel");
}

}  // namespace jaxib
