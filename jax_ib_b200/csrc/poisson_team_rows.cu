// K4 / K6, second generation ("team" row kernels): the packed-real FFT along y fused with the divergence
// (K4, finite_differences.py:139-146 + pressure.py:94-108) and with gradient / projection / pressure update
// (K6, pressure.py:75-89).  Same mathematics as poisson_fast.cu; the data movement follows what ncu showed
// for the first generation (profiles/r1i, r2d: 43 % issue utilisation at 20 % occupancy -- 128-thread CTAs
// at 122-128 registers, one row at a time, every pass fenced by a CTA barrier, one row of prefetch):
//
//   * ONE TEAM of T = N/16 threads per row (N = ny/2 complex points; 64 threads at ny = 2048), 1024 threads
//     = 1024/T teams per CTA, each team on its own row with NAMED barriers of T threads.  At 2048^2 the whole
//     problem is one wave of 128 (K4) / 137 (K6) CTAs: every row of the grid is in flight at once;
//   * no staging: the first pass reads u, v straight from global memory into registers (each team reads
//     512-byte contiguous runs), computes the divergence and transforms; the last inverse pass leaves q in
//     registers / shared memory for the projection.  Shared memory holds only the FFT row of each team
//     (8.7 KB at ny = 2048) and the twiddle tables (one bulk asynchronous copy per CTA, UBLKCP + mbarrier);
//   * 16 complex values per thread and pass, padded layout p + (p >> 4) + (p >> PS2): conflict-free passes,
//     every address = per-thread base + compile-time constant (tools/fft_model.py checks both);
//   * K6: the u-correction of row i needs q of row i+1, which lives in the NEXT team's buffer: a CTA of NT
//     teams transforms NT consecutive rows and finalises the first NT-1 of them after one CTA barrier (the
//     first generation carried q from row to row inside a CTA and could not run rows concurrently).
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "fft.cuh"
#include "poisson_fast.cuh"

namespace jaxib {
namespace {

constexpr unsigned FULLW = 0xffffffffu;

template <int R, bool INV>
__device__ __forceinline__ void dft_r(float2* v) {
  if (INV) {
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = cswap(v[r]);
  }
  Dft<R>::run(v);
  if (INV) {
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = cswap(v[r]);
  }
}

// Barrier of one row team.  Only 15 named hardware barriers exist besides barrier 0, so teams smaller than 128
// threads share one barrier per 128 threads (a one-warp team just uses __syncwarp).  Teams that share a barrier
// must execute the same barrier sequence: idle teams run through the barriers with their work predicated off.
template <class G>
__device__ __forceinline__ void team_bar(int team) {
  if (G::T == 32) {
    __syncwarp();
  } else {
    constexpr int TPB = G::T >= 128 ? 1 : 128 / G::T;  // teams per barrier
    asm volatile("bar.sync %0, %1;" ::"r"(team / TPB + 1), "r"(G::T * TPB) : "memory");
  }
}

template <int R1, int R2, int R3, int PS2>
struct RowTeam {
  static constexpr int N = R1 * R2 * R3;  // complex length ny/2
  static constexpr int T = N / 16;        // threads per row team
  static constexpr int NT = 1024 / T;     // teams per CTA at most (the launcher may use fewer: small grids)
  static constexpr int SUB1 = N / R1, SUB2 = SUB1 / R2;
  static constexpr int NBA = 16 / R1, NBB = 16 / R2, NBC = 16 / R3;
  __host__ __device__ static constexpr int pad(int p) { return p + (p >> 4) + (p >> PS2); }
  static constexpr int BUF = (pad(N) + 3) & ~1;  // floats2 per team buffer (even: 16-byte aligned buffers)
  static constexpr int TA = (R1 - 1) * SUB1, TB = (R2 - 1) * SUB2, TU = N / 2 + 1;
  static constexpr int TAB = (TA + TB + TU + 1) & ~1;  // table entries (even: bulk copies move 16-byte units)
  __host__ static size_t smem(int nt) { return (size_t)(TAB + nt * BUF) * sizeof(float2); }
  // position of frequency k after the forward DIF (digit reversal of the radix triple)
  __device__ static __forceinline__ int pos(int k) {
    return (k % R1) * SUB1 + ((k / R1) % R2) * SUB2 + k / (R1 * R2);
  }
  static_assert(T >= 32 && 1024 % T == 0, "whole warps per team");
};

// the CTA's twiddle tables [TA | TB | TU]: one bulk copy, completion on an mbarrier (phase 0)
template <class G>
__device__ __forceinline__ void tables_issue(float2* dst, const float2* src, unsigned bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  constexpr unsigned bytes = (unsigned)(G::TAB * sizeof(float2));
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tables_wait(unsigned bar) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(bar)
                 : "memory");
}

// inner passes on one team's shared-memory row
template <class G, int R2, int R3, int R1, bool INV>
__device__ __forceinline__ void team_inner(float2* buf, const float2* tb, int b, int team, bool live) {
  constexpr int T = G::T, SUB1 = G::SUB1, SUB2 = G::SUB2;
  auto passB = [&]() {
    if (!live) return;
#pragma unroll
    for (int j = 0; j < G::NBB; ++j) {
      const int bf = b + T * j;
      const int blk = bf / SUB2, b2 = bf - blk * SUB2;
      float2* const e = buf + G::pad(blk * SUB1 + b2);
      float2 v[R2];
#pragma unroll
      for (int r = 0; r < R2; ++r) v[r] = e[G::pad(SUB2 * r)];
      if (!INV) {
        Dft<R2>::run(v);
#pragma unroll
        for (int q = 1; q < R2; ++q) v[q] = cmul(v[q], tb[(q - 1) * SUB2 + b2]);
      } else {
#pragma unroll
        for (int q = 1; q < R2; ++q) v[q] = cmulc(v[q], tb[(q - 1) * SUB2 + b2]);
        dft_r<R2, true>(v);
      }
#pragma unroll
      for (int r = 0; r < R2; ++r) e[G::pad(SUB2 * r)] = v[r];
    }
  };
  auto passC = [&]() {
    if (!live) return;
#pragma unroll
    for (int j = 0; j < G::NBC; ++j) {
      float2* const e = buf + G::pad((b + T * j) * R3);
      float2 v[R3];
#pragma unroll
      for (int r = 0; r < R3; ++r) v[r] = e[G::pad(r)];
      dft_r<R3, INV>(v);
#pragma unroll
      for (int r = 0; r < R3; ++r) e[G::pad(r)] = v[r];
    }
  };
  if (!INV) {
    passB();
    team_bar<G>(team);
    passC();
    team_bar<G>(team);
  } else {
    passC();
    team_bar<G>(team);
    passB();
    team_bar<G>(team);
  }
}

// ------------------------------------------------------------------------------------------------ K4
template <int R1, int R2, int R3, int PS2>
__global__ void __launch_bounds__(1024, 1)
rows_fwd_team_kernel(FastRowParams P, const float2* __restrict__ tables, const float* __restrict__ u,
                     const float* __restrict__ v, float2* __restrict__ spec) {
  using G = RowTeam<R1, R2, R3, PS2>;
  constexpr int N = G::N, T = G::T, SUB1 = G::SUB1;
  extern __shared__ __align__(16) float2 smr[];
  __shared__ __align__(8) unsigned long long tbar;
  float2* const ta = smr;
  float2* const tb = ta + G::TA;
  float2* const tu = tb + G::TB;
  const int t = threadIdx.x;
  const int team = t / T, b = t - team * T, lane = t & 31;
  float2* const buf = smr + G::TAB + team * G::BUF;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&tbar);
  pdl_trigger();
  if (t == 0) tables_issue<G>(smr, tables, bar);  // independent of the previous kernel
  const GridF& g = P.g;
  const int ny = g.ny;
  const int i = blockIdx.x * (blockDim.x / T) + team;
  const bool live = i < g.nx;  // uniform over the team (whole warps)
  const size_t plane = (size_t)g.nx * ny;
  u += blockIdx.y * plane;
  v += blockIdx.y * plane;
  spec += (size_t)blockIdx.y * g.nx * P.sp;
  __syncthreads();  // barrier initialisation visible (before the dependency waits: teams wait for different things)
  // Dependency: a row far from every body only needs the predictor (RowDeps in common.cuh), everything else waits
  // for the previous kernel as usual.
  bool far = false;
  if (P.deps.k1_counter != nullptr && live) {
    bool hit = false;
    for (int b2 = lane; b2 < P.deps.n_bodies; b2 += 32) {
      const float xc = __ldg(P.deps.body_state + (size_t)b2 * JAXIB_BODY_STATE_STRIDE + 2);
      const int ic = (int)floorf(__fsub_rn(__fdiv_rn(xc, g.dx), 1.0f));  // u-cell row of the centre (ib.cu: cell_index)
      hit |= i >= ic - P.deps.half_extent - 1 && i <= ic + P.deps.half_extent + 2;
    }
    far = !__any_sync(FULLW, hit);
  }
  if (far) {
    bool done = true;
    if (lane == 0) {
      const long long t0 = clock64();
      unsigned seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(P.deps.k1_counter) : "memory");
        done = (int)(seen - P.deps.k1_target) >= 0;
      } while (!done && clock64() - t0 < 600000);  // ~0.3 ms: far beyond any predictor; then the ordinary wait
    }
    done = __shfl_sync(FULLW, done ? 1 : 0, 0) != 0;
    __syncwarp();
    if (!done) pdl_wait();
  } else {
    pdl_wait();
  }
  // slab mode (nx_global > 0): the rows are open -- u row -1 sits next to the array (halo)
  const ptrdiff_t im1 = (i == 0 && g.nx_global == 0) ? g.nx - 1 : i - 1;
  const float2* __restrict__ ur = reinterpret_cast<const float2*>(u + (size_t)i * ny);
  const float2* __restrict__ um = reinterpret_cast<const float2*>(u + im1 * ny);
  const float* __restrict__ vrow = v + (size_t)i * ny;
  const float2* __restrict__ vr = reinterpret_cast<const float2*>(vrow);
  bool have_tables = false;
#pragma unroll
  for (int j = 0; j < G::NBA; ++j) {  // pass A: global -> divergence -> registers -> shared
    if (!live) break;
    const int bA = b + T * j;
    float2 z[R1];
    // loads in chunks of 8 rows-elements: (a, c) then (d, edge) -- at most 32 registers of loads in flight
#pragma unroll
    for (int h = 0; h < R1; h += 8) {
      float2 a[8], c[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) { a[r] = ur[bA + SUB1 * (h + r)]; c[r] = um[bA + SUB1 * (h + r)]; }
#pragma unroll
      for (int r = 0; r < 8; ++r) z[h + r] = make_float2((a[r].x - c[r].x) * g.inv_dx, (a[r].y - c[r].y) * g.inv_dx);
    }
#pragma unroll
    for (int h = 0; h < R1; h += 8) {
      float2 d[8];
      float edge[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int n = bA + SUB1 * (h + r);
        d[r] = vr[n];
        edge[r] = 0.0f;
        if (lane == 0) edge[r] = vrow[n > 0 ? 2 * n - 1 : ny - 1];  // v[i][2n-1] of the neighbouring warp (periodic)
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float vm = __shfl_up_sync(FULLW, d[r].y, 1);
        if (lane == 0) vm = edge[r];
        z[h + r].x = fmaf(d[r].x - vm, g.inv_dy, z[h + r].x);
        z[h + r].y = fmaf(d[r].y - d[r].x, g.inv_dy, z[h + r].y);
      }
    }
    if (!have_tables) { tables_wait(bar); have_tables = true; }
    Dft<R1>::run(z);
    float2* const dst = buf + G::pad(bA);
#pragma unroll
    for (int q = 0; q < R1; ++q) {
      if (q > 0) z[q] = cmul(z[q], ta[(q - 1) * SUB1 + bA]);
      dst[G::pad(SUB1 * q)] = z[q];
    }
  }
  team_bar<G>(team);
  team_inner<G, R2, R3, R1, false>(buf, tb, b, team, live);
  if (!live) return;
  // untangle: X[k] = E + W^k O, X[N-k] = conj(E - W^k O)
  float2* const out = spec + (size_t)i * P.sp;
  const bool scatter = P.sc.cpr != 0;  // slab mode: column k lives in the buffer of the slab that owns it
  auto put = [&](int k, float2 val) {
    if (scatter) *spec_addr(P.sc, i, k) = val;
    else out[k] = val;
  };
#pragma unroll
  for (int m = 0; m < (N / 2) / T; ++m) {
    const int k = b + T * m;
    if (k == 0) {
      const float2 zz = buf[0];
      put(0, make_float2(zz.x + zz.y, 0.0f));
      put(N, make_float2(zz.x - zz.y, 0.0f));
    } else {
      const int km = N - k;
      const float2 zk = buf[G::pad(G::pos(k))], zm = cconj(buf[G::pad(G::pos(km))]);
      const float2 e = cscale(cadd(zk, zm), 0.5f);
      const float2 o = cscale(mul_mi(csub(zk, zm)), 0.5f);
      const float2 tt = cmul(tu[k], o);
      put(k, cadd(e, tt));
      put(km, cconj(csub(e, tt)));
    }
  }
  if (b == 0) {  // k = N/2 is its own mirror
    const int k = N / 2;
    const float2 zk = buf[G::pad(G::pos(k))], zm = cconj(zk);
    const float2 e = cscale(cadd(zk, zm), 0.5f);
    const float2 o = cscale(mul_mi(csub(zk, zm)), 0.5f);
    put(k, cadd(e, cmul(tu[k], o)));
  }
  if (b < 3) put(N + 1 + b, make_float2(0.0f, 0.0f));  // padding columns feed the column kernel: keep them finite
}

// ------------------------------------------------------------------------------------------------ K6
template <int R1, int R2, int R3, int PS2, bool PROJECT>
__global__ void __launch_bounds__(1024, 1)
rows_inv_team_kernel(FastRowParams P, const float2* __restrict__ tables, const float2* __restrict__ spec, const float* u,
                     const float* v, const float* p, float* u_out, float* v_out, float* p_out) {
  using G = RowTeam<R1, R2, R3, PS2>;
  constexpr int N = G::N, T = G::T, SUB1 = G::SUB1;
  const int NT = blockDim.x / T;               // teams in this CTA
  const int ROWS = PROJECT ? NT - 1 : NT;      // rows finalised per CTA (the last team carries the row after them)
  extern __shared__ __align__(16) float2 smr[];
  __shared__ __align__(8) unsigned long long tbar;
  float2* const ta = smr;
  float2* const tb = ta + G::TA;
  float2* const tu = tb + G::TB;
  const int t = threadIdx.x;
  const int team = t / T, b = t - team * T;
  float2* const buf = smr + G::TAB + team * G::BUF;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&tbar);
  pdl_trigger();
  if (t == 0) tables_issue<G>(smr, tables, bar);
  const GridF& g = P.g;
  const int ny = g.ny;
  const size_t plane = (size_t)g.nx * ny;
  spec += (size_t)blockIdx.y * g.nx * P.sp;
  if (PROJECT) { u += blockIdx.y * plane; v += blockIdx.y * plane; p += blockIdx.y * plane; u_out += blockIdx.y * plane; v_out += blockIdx.y * plane; }
  p_out += blockIdx.y * plane;
  const int i0 = blockIdx.x * ROWS;
  const int ii = i0 + team;
  const int i_last = min(i0 + ROWS, g.nx);                 // first row NOT finalised by this CTA
  const bool live = PROJECT ? ii <= i_last : ii < g.nx;    // this team transforms a row (uniform over the team)
  // slab mode (nx_global > 0): spectrum row nx exists (halo); otherwise the row after the last one wraps
  const int irow = (ii >= g.nx && g.nx_global == 0) ? ii - g.nx : ii;
  pdl_wait();
  __syncthreads();  // barrier initialisation visible
  if (live) {
    const float2* __restrict__ in = spec + (size_t)irow * P.sp;
    {  // tangle: spectrum row -> packed complex sequence at its digit-reversed positions
      float2 xk[(N / 2) / T], xm[(N / 2) / T];
#pragma unroll
      for (int m = 0; m < (N / 2) / T; ++m) {
        const int k = b + T * m;
        xk[m] = in[k];
        xm[m] = in[N - k];
      }
      float2 xh = make_float2(0.f, 0.f);
      if (b == 0) xh = in[N / 2];
      tables_wait(bar);
#pragma unroll
      for (int m = 0; m < (N / 2) / T; ++m) {
        const int k = b + T * m;
        if (k == 0) {
          const float x0 = xk[m].x, xn = xm[m].x;  // X[0], X[N]: both real
          buf[0] = make_float2(0.5f * (x0 + xn), 0.5f * (x0 - xn));
        } else {
          const float2 a = xk[m], c = cconj(xm[m]);
          const float2 e = cscale(cadd(a, c), 0.5f);
          const float2 o = cmulc(cscale(csub(a, c), 0.5f), tu[k]);
          buf[G::pad(G::pos(k))] = cadd(e, mul_pi(o));
          buf[G::pad(G::pos(N - k))] = cadd(cconj(e), mul_pi(cconj(o)));
        }
      }
      if (b == 0) {
        const float2 a = xh, c = cconj(xh);
        const float2 e = cscale(cadd(a, c), 0.5f);
        const float2 o = cmulc(cscale(csub(a, c), 0.5f), tu[N / 2]);
        buf[G::pad(G::pos(N / 2))] = cadd(e, mul_pi(o));
      }
    }
  }
  team_bar<G>(team);
  team_inner<G, R2, R3, R1, true>(buf, tb, b, team, live);
  if (live) {
#pragma unroll
    for (int j = 0; j < G::NBA; ++j) {  // pass A inverse; q[i][2n], q[i][2n+1] stay at (padded) index n = bA + SUB1 r
      const int bA = b + T * j;
      float2* const e = buf + G::pad(bA);
      float2 z[R1];
#pragma unroll
      for (int r = 0; r < R1; ++r) {
        z[r] = e[G::pad(SUB1 * r)];
        if (r > 0) z[r] = cmulc(z[r], ta[(r - 1) * SUB1 + bA]);
      }
      dft_r<R1, true>(z);
      if (!PROJECT) {
        float2* qo = reinterpret_cast<float2*>(p_out + (size_t)irow * ny);
#pragma unroll
        for (int r = 0; r < R1; ++r) qo[bA + SUB1 * r] = cscale(z[r], P.scale);
      } else {
#pragma unroll
        for (int r = 0; r < R1; ++r) e[G::pad(SUB1 * r)] = cscale(z[r], P.scale);
      }
    }
  }
  if (!PROJECT) return;
  __syncthreads();  // every team's q row is complete
  if (team >= ROWS || ii >= g.nx) return;
  const float2* const q0row = buf;
  const float2* const q1row = buf + G::BUF;  // the next team's row = q[i+1]
  const size_t ro = (size_t)ii * ny;
  const float2* u2 = reinterpret_cast<const float2*>(u + ro);
  const float2* v2 = reinterpret_cast<const float2*>(v + ro);
  const float2* p2 = reinterpret_cast<const float2*>(p + ro);
  float2* uo = reinterpret_cast<float2*>(u_out + ro);
  float2* vo = reinterpret_cast<float2*>(v_out + ro);
  float2* po = reinterpret_cast<float2*>(p_out + ro);
  constexpr int CH = 4;  // elements per thread in flight
#pragma unroll 1
  for (int m0 = 0; m0 < N / T; m0 += CH) {
    // all loads first: the outputs may alias the inputs (in-place projection)
    float2 ua[CH], va[CH], pa[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int n = b + T * (m0 + c);
      ua[c] = u2[n];
      va[c] = v2[n];
      pa[c] = p2[n];
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int n = b + T * (m0 + c);
      const float2 q0 = q0row[G::pad(n)], q1 = q1row[G::pad(n)];
      const float qn = q0row[G::pad(n + 1 < N ? n + 1 : 0)].x;  // q[i][2n+2] (periodic)
      float2 uu = ua[c], vv = va[c];
      uu.x -= (q1.x - q0.x) * g.inv_dx;
      uu.y -= (q1.y - q0.y) * g.inv_dx;
      vv.x -= (q0.y - q0.x) * g.inv_dy;
      vv.y -= (qn - q0.y) * g.inv_dy;
      const float2 pn = make_float2(q0.x + pa[c].x, q0.y + pa[c].y);
      uo[n] = uu;
      vo[n] = vv;
      po[n] = pn;
      if (P.ho.halo) {  // slab mode: boundary rows are also the neighbours' halo rows (NVLink stores)
        if (ii < P.ho.halo) {
          const size_t o = (size_t)ii * ny;
          reinterpret_cast<float2*>(P.ho.lo[0] + o)[n] = uu;
          reinterpret_cast<float2*>(P.ho.lo[1] + o)[n] = vv;
          reinterpret_cast<float2*>(P.ho.lo[2] + o)[n] = pn;
        }
        if (ii >= g.nx - P.ho.halo) {
          const size_t o = (size_t)(ii - (g.nx - P.ho.halo)) * ny;
          reinterpret_cast<float2*>(P.ho.hi[0] + o)[n] = uu;
          reinterpret_cast<float2*>(P.ho.hi[1] + o)[n] = vv;
          reinterpret_cast<float2*>(P.ho.hi[2] + o)[n] = pn;
        }
      }
    }
  }
}

// K6, marching variant: one CTA per CHUNK of L consecutive rows, walked from the chunk's last row to its first in
// batches of NTB rows (the first batch also transforms the row after the chunk: q of that row feeds the u-correction
// of the chunk's last row; later batches keep q of the row above them in an extra buffer).  L is chosen so that the
// grid is one wave: with long rows a CTA only holds NT = 4 teams and the plain kernel finalises 3 rows per CTA
// -- a 1024-row slab of 8192 columns is 342 CTAs = 3 waves for 2.3 waves of work, while 147 chunks of 7 rows are
// 2 batches of 4 transforms each.  Measured: no gain (see launch_rows_inv_team_t) -- kept off by default, tested bitwise.
template <int R1, int R2, int R3, int PS2>
__global__ void __launch_bounds__(1024, 1)
rows_inv_march_kernel(FastRowParams P, const float2* __restrict__ tables, const float2* __restrict__ spec, const float* u,
                      const float* v, const float* p, float* u_out, float* v_out, float* p_out, int L) {
  using G = RowTeam<R1, R2, R3, PS2>;
  constexpr int N = G::N, T = G::T, SUB1 = G::SUB1;
  const int NTB = blockDim.x / T;
  extern __shared__ __align__(16) float2 smr[];
  __shared__ __align__(8) unsigned long long tbar;
  float2* const ta = smr;
  float2* const tb = ta + G::TA;
  float2* const tu = tb + G::TB;
  const int t = threadIdx.x;
  const int team = t / T, b = t - team * T;
  float2* const buf = smr + G::TAB + team * G::BUF;
  float2* const qsave = smr + G::TAB + NTB * G::BUF;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&tbar);
  pdl_trigger();
  if (t == 0) tables_issue<G>(smr, tables, bar);
  const GridF& g = P.g;
  const int ny = g.ny;
  const size_t plane = (size_t)g.nx * ny;
  spec += (size_t)blockIdx.y * g.nx * P.sp;
  u += blockIdx.y * plane; v += blockIdx.y * plane; p += blockIdx.y * plane;
  u_out += blockIdx.y * plane; v_out += blockIdx.y * plane; p_out += blockIdx.y * plane;
  const int i0 = blockIdx.x * L;
  const int Lc = min(L, g.nx - i0);
  pdl_wait();
  __syncthreads();  // barrier initialisation visible
  bool have_tables = false;
  for (int jhi = Lc; jhi >= 0; jhi -= NTB) {
    const int jbase = jhi - NTB + 1;
    const int j = jbase + team;  // this team's row of the chunk (Lc: the row after it)
    const bool live = j >= 0;    // uniform over the team
    const bool more = jbase > 0;
    if (live) {
      // slab mode (nx_global > 0): spectrum row nx exists (halo); otherwise the row after the last one wraps
      const int ii = i0 + j;
      const int irow = (ii >= g.nx && g.nx_global == 0) ? ii - g.nx : ii;
      const float2* __restrict__ in = spec + (size_t)irow * P.sp;
      float2 xk[(N / 2) / T], xm[(N / 2) / T];
#pragma unroll
      for (int m = 0; m < (N / 2) / T; ++m) {
        const int k = b + T * m;
        xk[m] = in[k];
        xm[m] = in[N - k];
      }
      float2 xh = make_float2(0.f, 0.f);
      if (b == 0) xh = in[N / 2];
      if (!have_tables) { tables_wait(bar); have_tables = true; }
#pragma unroll
      for (int m = 0; m < (N / 2) / T; ++m) {
        const int k = b + T * m;
        if (k == 0) {
          const float x0 = xk[m].x, xn = xm[m].x;  // X[0], X[N]: both real
          buf[0] = make_float2(0.5f * (x0 + xn), 0.5f * (x0 - xn));
        } else {
          const float2 a = xk[m], c = cconj(xm[m]);
          const float2 e = cscale(cadd(a, c), 0.5f);
          const float2 o = cmulc(cscale(csub(a, c), 0.5f), tu[k]);
          buf[G::pad(G::pos(k))] = cadd(e, mul_pi(o));
          buf[G::pad(G::pos(N - k))] = cadd(cconj(e), mul_pi(cconj(o)));
        }
      }
      if (b == 0) {
        const float2 a = xh, c = cconj(xh);
        const float2 e = cscale(cadd(a, c), 0.5f);
        const float2 o = cmulc(cscale(csub(a, c), 0.5f), tu[N / 2]);
        buf[G::pad(G::pos(N / 2))] = cadd(e, mul_pi(o));
      }
    }
    team_bar<G>(team);
    team_inner<G, R2, R3, R1, true>(buf, tb, b, team, live);
    if (live) {
#pragma unroll
      for (int jj = 0; jj < G::NBA; ++jj) {  // pass A inverse; q[i][2n], q[i][2n+1] stay at (padded) index n
        const int bA = b + T * jj;
        float2* const e = buf + G::pad(bA);
        float2 z[R1];
#pragma unroll
        for (int r = 0; r < R1; ++r) {
          z[r] = e[G::pad(SUB1 * r)];
          if (r > 0) z[r] = cmulc(z[r], ta[(r - 1) * SUB1 + bA]);
        }
        dft_r<R1, true>(z);
#pragma unroll
        for (int r = 0; r < R1; ++r) e[G::pad(SUB1 * r)] = cscale(z[r], P.scale);
      }
    }
    __syncthreads();  // every team's q row is complete
    if (live && j < Lc) {
      const int ii = i0 + j;
      const float2* const q0row = buf;
      const float2* const q1row = team + 1 < NTB ? buf + G::BUF : qsave;  // q[i+1]
      const size_t ro = (size_t)ii * ny;
      const float2* u2 = reinterpret_cast<const float2*>(u + ro);
      const float2* v2 = reinterpret_cast<const float2*>(v + ro);
      const float2* p2 = reinterpret_cast<const float2*>(p + ro);
      float2* uo = reinterpret_cast<float2*>(u_out + ro);
      float2* vo = reinterpret_cast<float2*>(v_out + ro);
      float2* po = reinterpret_cast<float2*>(p_out + ro);
      constexpr int CH = 4;
#pragma unroll 1
      for (int m0 = 0; m0 < N / T; m0 += CH) {
        float2 ua[CH], va[CH], pa[CH];  // all loads first: the outputs may alias the inputs
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int n = b + T * (m0 + c);
          ua[c] = u2[n];
          va[c] = v2[n];
          pa[c] = p2[n];
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int n = b + T * (m0 + c);
          const float2 q0 = q0row[G::pad(n)], q1 = q1row[G::pad(n)];
          const float qn = q0row[G::pad(n + 1 < N ? n + 1 : 0)].x;  // q[i][2n+2] (periodic)
          float2 uu = ua[c], vv = va[c];
          uu.x -= (q1.x - q0.x) * g.inv_dx;
          uu.y -= (q1.y - q0.y) * g.inv_dx;
          vv.x -= (q0.y - q0.x) * g.inv_dy;
          vv.y -= (qn - q0.y) * g.inv_dy;
          const float2 pn = make_float2(q0.x + pa[c].x, q0.y + pa[c].y);
          uo[n] = uu;
          vo[n] = vv;
          po[n] = pn;
          if (P.ho.halo) {  // slab mode: boundary rows are also the neighbours' halo rows (NVLink stores)
            if (ii < P.ho.halo) {
              const size_t o = (size_t)ii * ny;
              reinterpret_cast<float2*>(P.ho.lo[0] + o)[n] = uu;
              reinterpret_cast<float2*>(P.ho.lo[1] + o)[n] = vv;
              reinterpret_cast<float2*>(P.ho.lo[2] + o)[n] = pn;
            }
            if (ii >= g.nx - P.ho.halo) {
              const size_t o = (size_t)(ii - (g.nx - P.ho.halo)) * ny;
              reinterpret_cast<float2*>(P.ho.hi[0] + o)[n] = uu;
              reinterpret_cast<float2*>(P.ho.hi[1] + o)[n] = vv;
              reinterpret_cast<float2*>(P.ho.hi[2] + o)[n] = pn;
            }
          }
        }
      }
    }
    if (more) {  // keep q of the batch's lowest row (team 0) for the next batch's highest one
      __syncthreads();
      if (team == 0)
        for (int n = b; n < G::BUF; n += T) qsave[n] = buf[n];
      __syncthreads();
    }
  }
}

// Teams per CTA: the most that still gives every SM a CTA (small grids would otherwise run on a fraction of the
// machine: 1024 rows of ny = 1024 are 32 CTAs of 32 teams); `halo` = 1 when one team per CTA carries the extra row.
int teams_per_cta(int max_nt, int nx, int batch, int halo) {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const long long rows = (long long)nx * batch;
  int nt = max_nt;
  while (nt > 4 && rows / (nt - halo) < sms) nt /= 2;
  // a barrier is shared by 128 threads: keep whole barrier groups
  return nt;
}

template <int R1, int R2, int R3, int PS2>
int launch_rows_fwd_team_t(cudaStream_t s, const FastRowParams& P, const float2* tables, const float* u, const float* v,
                           float2* spec) {
  using G = RowTeam<R1, R2, R3, PS2>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(rows_fwd_team_kernel<R1, R2, R3, PS2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem(G::NT));
    configured = true;
  }
  int nt = teams_per_cta(G::NT, P.g.nx, P.g.batch, 0);
  if (P.deps.k1_counter != nullptr) {
    // early start: the CTAs must fit NEXT TO the IB kernels' CTAs (half the registers of an SM at most)
    static const int cap = getenv("JAXIB_K4_EARLY_THREADS") ? atoi(getenv("JAXIB_K4_EARLY_THREADS")) : 512;
    while (nt > 2 && nt * G::T > cap) nt /= 2;
  }
  dim3 grid((P.g.nx + nt - 1) / nt, P.g.batch);
  launch_pdl(8, rows_fwd_team_kernel<R1, R2, R3, PS2>, grid, dim3(nt * G::T), G::smem(nt), s, P, tables, u, v, spec);
  return check_launch("rows_fwd_team_kernel");
}

template <int R1, int R2, int R3, int PS2>
int launch_rows_inv_team_t(cudaStream_t s, const FastRowParams& P, const float2* tables, const float2* spec, const float* u,
                           const float* v, const float* p, float* uo, float* vo, float* po, bool project) {
  using G = RowTeam<R1, R2, R3, PS2>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(rows_inv_team_kernel<R1, R2, R3, PS2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem(G::NT));
    cudaFuncSetAttribute(rows_inv_team_kernel<R1, R2, R3, PS2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem(G::NT));
    configured = true;
  }
  const int nt = teams_per_cta(G::NT, P.g.nx, P.g.batch, project ? 1 : 0);
  const int rows = project ? nt - 1 : nt;
  if (project && P.g.batch == 1) {
    // marching variant (rows_inv_march_kernel): one wave of chunks.  Cost model in row transforms per SM:
    //   plain: waves * nt;  marching: ceil((L + 1) / nt) * nt with L = ceil(nx / SMs)
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const int big = (int)G::smem(G::NT + 1);
      cudaFuncSetAttribute(rows_inv_march_kernel<R1, R2, R3, PS2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    }
    const char* env_march = getenv("JAXIB_K6_MARCH");  // 0 / 1 forces the choice (A/B runs, tests)
    const int mode = env_march ? atoi(env_march) : -1;
    const int ctas = (P.g.nx + rows - 1) / rows;
    const int plain = ((ctas + sms - 1) / sms) * nt;
    const int L = (P.g.nx + sms - 1) / sms;
    const int march = ((L + 1 + nt - 1) / nt) * nt;
    const bool fits = G::smem(nt + 1) <= 227 * 1024;
    // Measured on B200 (profiles/r3h): NOT faster -- 63.3 vs 62.2 us on a 1024-row slab of 8192 columns, 1360 vs 1301 us
    // per step at 8192^2 on one GPU: the batches of a chunk serialise load / transform / finalise phases that the
    // independent CTAs of the plain kernel overlap across waves.  Off unless forced (JAXIB_K6_MARCH=1); the cost
    // model is kept for the record.
    (void)plain; (void)march;
    if (fits && mode == 1) {
      dim3 grid((P.g.nx + L - 1) / L, 1);
      launch_pdl(32, rows_inv_march_kernel<R1, R2, R3, PS2>, grid, dim3(nt * G::T), G::smem(nt + 1), s, P, tables, spec, u, v, p,
                 uo, vo, po, L);
      return check_launch("rows_inv_march_kernel");
    }
  }
  dim3 grid((P.g.nx + rows - 1) / rows, P.g.batch);
  if (project)
    launch_pdl(32, rows_inv_team_kernel<R1, R2, R3, PS2, true>, grid, dim3(nt * G::T), G::smem(nt), s, P, tables, spec, u, v, p, uo, vo, po);
  else
    launch_pdl(32, rows_inv_team_kernel<R1, R2, R3, PS2, false>, grid, dim3(nt * G::T), G::smem(nt), s, P, tables, spec, u, v, p, uo, vo, po);
  return check_launch("rows_inv_team_kernel");
}


// ================================================================================================ x-sweep
// Third generation of the periodic pressure solve: NO column transform.  After the row FFT along y every
// y-wavenumber k has its own 1-D periodic problem along x,
//     Q[i-1] - (2 + mu_k) Q[i] + Q[i+1] = dx^2 X[i],      mu_k = -ly_k dx^2 >= 0   (pressure.py:94-108: the
// eigenvalues lx + ly of the same second-difference operator, array_utils.py:168-173),
// whose Green's function is a two-sided geometric sequence: with rho_k in (0, 1) the root of
// rho + 1/rho = 2 + mu_k,
//     a[i] = X[i] + rho a[i-1]   (forward sweep),   b[i] = a[i] + rho b[i+1]   (backward sweep),   Q = -dx^2 rho b.
// Both sweeps are first-order recurrences ALONG x and elementwise ACROSS k, which is exactly how the row kernels
// already walk the grid: K4 runs the forward sweep over the rows of its chunk right after their FFTs, K6 the
// backward sweep right before its inverse FFTs.  Chunks start from zero; the true values differ from the
// chunk-local ones by rho^(r+1) times the value carried into the chunk, and the carried values of all chunks
// obey a short cyclic recurrence that a tiny kernel solves in double precision between K4 and K6
// (sweep_carry_kernel).  The k = 0 column (mu = 0, singular: the mean mode is projected out, as the
// pseudo-inverse does) is solved exactly by two prefix sums in that kernel.
//
// Against the FFT x FFT solve this removes the whole column kernel (8 B per cell of HBM traffic, ~20 % of the
// step at 2048^2), works for any nx, and -- decisive for the slab decomposition -- the x-direction couples the
// slabs only through one carried value per chunk and wavenumber instead of two full transposes of the spectrum.
// Accuracy: tools/sweep_model.py reproduces the arithmetic in NumPy float32: rel-L2 1.9e-7 against the float64
// solution at 2048^2 (float32 FFT solve: 1.7e-7).
//
// Row image: the chunk-local forward sweep is stored in the SLOT order of the team buffers (frequency k at
// pad(pos(k)), the Nyquist column in the extra slot pad(N), padding holes unused): K4 untangles in place, the
// sweeps are elementwise, so consecutive threads move consecutive 8-byte words of a row in both kernels.

constexpr int kSweepSlots = 5;   // image entries per thread at most (>= 4 teams per CTA)
constexpr int kSweepMaxTeams = 16;  // teams (= rows in flight) per CTA at most: the sweeps hold one value per row in registers

template <int R1, int R2, int R3, int PS2>
__global__ void __launch_bounds__(1024, 1)
rows_fwd_sweep_kernel(SweepParams P, const float2* __restrict__ tables, const float* __restrict__ u,
                      const float* __restrict__ v) {
  using G = RowTeam<R1, R2, R3, PS2>;
  constexpr int N = G::N, T = G::T, SUB1 = G::SUB1;
  constexpr int NTM = G::NT < kSweepMaxTeams ? G::NT : kSweepMaxTeams;
  extern __shared__ __align__(16) float2 smr[];
  __shared__ __align__(8) unsigned long long tbar;
  float2* const ta = smr;
  float2* const tb = ta + G::TA;
  float2* const tu = tb + G::TB;
  float2* const bufs = smr + G::TAB;
  const int t = threadIdx.x, nthreads = blockDim.x, NTB = nthreads / T;
  const int team = t / T, b = t - team * T, lane = t & 31;
  float2* const buf = bufs + team * G::BUF;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&tbar);
  pdl_trigger();
  if (t == 0) tables_issue<G>(smr, tables, bar);  // independent of the previous kernel
  const GridF& g = P.g;
  const int ny = g.ny;
  const int i0 = blockIdx.x * P.L;
  const int Lc = min(P.L, g.nx - i0);
  const size_t plane = (size_t)g.nx * ny;
  u += blockIdx.y * plane;
  v += blockIdx.y * plane;
  float2* const arow = P.a + ((size_t)blockIdx.y * g.nx + i0) * P.img;
  float2* const agg = P.agg + ((size_t)blockIdx.y * P.C + blockIdx.x) * P.img;
  pdl_wait();
  __syncthreads();  // barrier initialisation visible
  bool have_tables = false;
  for (int m0 = 0; m0 < Lc; m0 += NTB) {
    const int i = i0 + m0 + team;
    const bool live = m0 + team < Lc;  // uniform over the team
    if (live) {
      const ptrdiff_t im1 = i == 0 ? g.nx - 1 : i - 1;
      const float2* __restrict__ ur = reinterpret_cast<const float2*>(u + (size_t)i * ny);
      const float2* __restrict__ um = reinterpret_cast<const float2*>(u + im1 * ny);
      const float* __restrict__ vrow = v + (size_t)i * ny;
      const float2* __restrict__ vr = reinterpret_cast<const float2*>(vrow);
#pragma unroll
      for (int j = 0; j < G::NBA; ++j) {  // pass A: global -> divergence -> registers -> shared (as rows_fwd_team_kernel)
        const int bA = b + T * j;
        float2 z[R1];
#pragma unroll
        for (int h = 0; h < R1; h += 8) {
          float2 a[8], c[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) { a[r] = ur[bA + SUB1 * (h + r)]; c[r] = um[bA + SUB1 * (h + r)]; }
#pragma unroll
          for (int r = 0; r < 8; ++r) z[h + r] = make_float2((a[r].x - c[r].x) * g.inv_dx, (a[r].y - c[r].y) * g.inv_dx);
        }
#pragma unroll
        for (int h = 0; h < R1; h += 8) {
          float2 d[8];
          float edge[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int n = bA + SUB1 * (h + r);
            d[r] = vr[n];
            edge[r] = 0.0f;
            if (lane == 0) edge[r] = vrow[n > 0 ? 2 * n - 1 : ny - 1];
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            float vm = __shfl_up_sync(FULLW, d[r].y, 1);
            if (lane == 0) vm = edge[r];
            z[h + r].x = fmaf(d[r].x - vm, g.inv_dy, z[h + r].x);
            z[h + r].y = fmaf(d[r].y - d[r].x, g.inv_dy, z[h + r].y);
          }
        }
        if (!have_tables) { tables_wait(bar); have_tables = true; }
        Dft<R1>::run(z);
        float2* const dst = buf + G::pad(bA);
#pragma unroll
        for (int q = 0; q < R1; ++q) {
          if (q > 0) z[q] = cmul(z[q], ta[(q - 1) * SUB1 + bA]);
          dst[G::pad(SUB1 * q)] = z[q];
        }
      }
    }
    team_bar<G>(team);
    team_inner<G, R2, R3, R1, false>(buf, tb, b, team, live);
    if (live) {  // untangle IN PLACE: X[k] to the slot of Z[k], X[N] to the extra slot
#pragma unroll
      for (int m = 0; m < (N / 2) / T; ++m) {
        const int k = b + T * m;
        if (k == 0) {
          const float2 zz = buf[0];
          buf[0] = make_float2(zz.x + zz.y, 0.0f);
          buf[G::pad(N)] = make_float2(zz.x - zz.y, 0.0f);
        } else {
          const int pk = G::pad(G::pos(k)), pm = G::pad(G::pos(N - k));
          const float2 zk = buf[pk], zm = cconj(buf[pm]);
          const float2 e = cscale(cadd(zk, zm), 0.5f);
          const float2 o = cscale(mul_mi(csub(zk, zm)), 0.5f);
          const float2 tt = cmul(tu[k], o);
          buf[pk] = cadd(e, tt);
          buf[pm] = cconj(csub(e, tt));
        }
      }
      if (b == 0) {  // k = N/2 is its own mirror
        const int pk = G::pad(G::pos(N / 2));
        const float2 zk = buf[pk], zm = cconj(zk);
        const float2 e = cscale(cadd(zk, zm), 0.5f);
        const float2 o = cscale(mul_mi(csub(zk, zm)), 0.5f);
        buf[pk] = cadd(e, cmul(tu[N / 2], o));
      }
    }
    __syncthreads();  // every row of the batch is untangled
    // forward sweep over the rows of the batch, elementwise over the image
    const int nr = min(NTB, Lc - m0);
    for (int s = 0; s < kSweepSlots; ++s) {
      const int e = t + s * nthreads;
      if (e >= G::BUF) break;
      const float rho = __ldg(P.rho + e);
      if (rho < 0.0f) continue;  // padding hole
      float2 ap = make_float2(0.f, 0.f), S = make_float2(0.f, 0.f);
      float w = 1.0f;
      if (m0 > 0) {  // state of the previous batch: its last row / the running sum (written by this thread)
        ap = arow[(size_t)(m0 - 1) * P.img + e];
        S = agg[e];
        w = exp2f((float)m0 * __ldg(P.lg + e));
      }
      float2* const col = bufs + e;
#pragma unroll
      for (int r = 0; r < NTM; ++r)
        if (r < nr) {
          const float2 x = col[r * G::BUF];
          ap.x = fmaf(rho, ap.x, x.x);
          ap.y = fmaf(rho, ap.y, x.y);
          col[r * G::BUF] = ap;
          arow[(size_t)(m0 + r) * P.img + e] = ap;
        }
      float2 h = make_float2(0.f, 0.f);  // sum_r rho^r a[r] of the batch (Horner from the last row)
#pragma unroll
      for (int r = NTM - 1; r >= 0; --r)
        if (r < nr) {
          const float2 x = col[r * G::BUF];
          h.x = fmaf(rho, h.x, x.x);
          h.y = fmaf(rho, h.y, x.y);
        }
      agg[e] = make_float2(fmaf(w, h.x, S.x), fmaf(w, h.y, S.y));
    }
    if (m0 + NTB < Lc) __syncthreads();  // the buffers are rewritten by the next batch
  }
}

// Carried values of the sweeps.  With e_c = the chunk-local forward value of chunk c's last row, P_c = rho^Lc:
//   forward   Aout[c] = e_c + P_c Ain[c],  Ain[c] = Aout[c-1]  (cyclic)
//   backward  Bout[c] = T_c + P_c Bin[c],  Bin[c] = Bout[c+1]  (cyclic),  T_c = S_c + Ain[c] rho (1 - rho^2Lc)/(1 - rho^2)
// (S_c = sum_r rho^r a_local[r]; the second term is the same sum over the carried part rho^(r+1) Ain).
// CTA = 32 image entries x 16 segments of consecutive chunks: a thread composes the affine maps of its segment,
// the 16 segment maps are closed cyclically from shared memory (geometric factor 1 / (1 - rho^nx)), and the
// thread then walks its chunks again with the incoming value.  All in double.  The last CTA of the grid solves
// the k = 0 column: Q[i-1] - 2 Q[i] + Q[i+1] = dx^2 (X[i] - mean X), mean Q = 0, by two prefix sums.
constexpr int kCarrySeg = 16;
constexpr int kCarryThreads = 32 * kCarrySeg;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLW, v, o);
  return v;
}
// sum over the CTA (every thread gets it); red: kCarryThreads / 32 doubles
__device__ __forceinline__ double block_sum(double v, double* red, int t) {
  v = warp_sum_d(v);
  __syncthreads();
  if ((t & 31) == 0) red[t >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < kCarryThreads / 32; ++w) s += red[w];
  return s;
}

// in-place scan of q[0..n) (inclusive or exclusive) by the whole CTA: every thread scans a contiguous piece, the
// piece totals are scanned by warp shuffles (two levels), the offsets are added back
__device__ void block_scan(double* q, int n, bool inclusive, double* red, int t) {
  const int per = (n + kCarryThreads - 1) / kCarryThreads;
  const int lo = min(t * per, n), hi = min(lo + per, n);
  __syncthreads();  // q was written with another thread mapping
  double s = 0.0;
  for (int i = lo; i < hi; ++i) s += q[i];
  double inc = s;  // inclusive scan of the piece totals within the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double y = __shfl_up_sync(FULLW, inc, o);
    if ((t & 31) >= o) inc += y;
  }
  __syncthreads();
  if ((t & 31) == 31) red[t >> 5] = inc;
  __syncthreads();
  double acc = inc - s;  // exclusive within the warp
  for (int w = 0; w < (t >> 5); ++w) acc += red[w];
  for (int i = lo; i < hi; ++i) {
    const double x = q[i];
    if (inclusive) { acc += x; q[i] = acc; } else { q[i] = acc; acc += x; }
  }
  __syncthreads();
}

// CPS: chunks per segment held in registers (the loads of a thread's chunks are issued together: three rounds of
// L2 latency in the whole kernel instead of one per chunk)
template <int CPS>
__global__ void __launch_bounds__(kCarryThreads)
sweep_carry_kernel(SweepParams P) {
  extern __shared__ double csm[];  // k = 0 CTA: nx doubles; others unused
  __shared__ double segE[2][kCarrySeg][32], segP[kCarrySeg][32], red[kCarryThreads / 32];
  pdl_trigger();
  const int t = threadIdx.x;
  const GridF& g = P.g;
  const int nx = g.nx, C = P.C, L = P.L;
  const size_t img = P.img;
  float2* const a = P.a + (size_t)blockIdx.y * nx * img;
  const float2* const agg = P.agg + (size_t)blockIdx.y * C * img;
  float2* const cA = P.cA + (size_t)blockIdx.y * C * img;
  float2* const cB = P.cB + (size_t)blockIdx.y * C * img;
  if (blockIdx.x == gridDim.x - 1) {  // ---- the k = 0 column (slot 0; real)
    pdl_wait();
    double sum = 0.0;
    for (int i = t; i < nx; i += kCarryThreads) { const double x = a[(size_t)i * img].x; csm[i] = x; sum += x; }
    const double mean = block_sum(sum, red, t) / nx;
    const double dx2 = (double)g.dx * (double)g.dx;
    for (int i = t; i < nx; i += kCarryThreads) csm[i] = (csm[i] - mean) * dx2;
    block_scan(csm, nx, true, red, t);  // s_i
    sum = 0.0;
    for (int i = t; i < nx; i += kCarryThreads) sum += csm[i];
    const double smean = block_sum(sum, red, t) / nx;
    for (int i = t; i < nx; i += kCarryThreads) csm[i] -= smean;  // d_i = Q[i+1] - Q[i]
    block_scan(csm, nx, false, red, t);  // Q_i - Q_0
    sum = 0.0;
    for (int i = t; i < nx; i += kCarryThreads) sum += csm[i];
    const double qmean = block_sum(sum, red, t) / nx;
    for (int i = t; i < nx; i += kCarryThreads) a[(size_t)i * img] = make_float2((float)(csm[i] - qmean), 0.0f);
    for (int c = t; c < C; c += kCarryThreads) {  // the row after chunk c (K6 differences q forward in x)
      int in = (c + 1) * L;
      if (in >= nx) in = 0;
      cB[(size_t)c * img] = make_float2((float)(csm[in] - qmean), 0.0f);
    }
    return;
  }
  const int el = t & 31, seg = t >> 5;
  const int e = blockIdx.x * 32 + el;
  const bool in_img = e < (int)img;
  const double rho = in_img ? P.rho_d[e] : -1.0;
  const bool act = rho > 0.0;  // holes (< 0) and the k = 0 slot (== 0) only walk through the barriers
  const int cps = (C + kCarrySeg - 1) / kCarrySeg;
  const int c0 = min(seg * cps, C), c1 = min(c0 + cps, C);
  const int Llast = nx - (C - 1) * L;
  const double PL = act ? pow(rho, (double)L) : 0.0, PLl = act ? pow(rho, (double)Llast) : 0.0;
  const double r2 = rho * rho;
  const double rG2 = act ? rho * (1.0 - PL * PL) / (1.0 - r2) : 0.0, rG2l = act ? rho * (1.0 - PLl * PLl) / (1.0 - r2) : 0.0;
  pdl_wait();
  float2 ec[CPS], Sc[CPS];  // ec: last row of the chunk-local forward sweep, later the carried-in value Ain
#pragma unroll
  for (int i = 0; i < CPS; ++i) {
    const int c = c0 + i;
    ec[i] = make_float2(0.f, 0.f);
    if (act && c < c1) ec[i] = a[(size_t)(min((c + 1) * L, nx) - 1) * img + e];
  }
#pragma unroll
  for (int i = 0; i < CPS; ++i) {
    const int c = c0 + i;
    Sc[i] = make_float2(0.f, 0.f);
    if (act && c < c1) Sc[i] = agg[(size_t)c * img + e];
  }
  // ---- forward: segment map
  double Ex = 0.0, Ey = 0.0, PP = 1.0;
#pragma unroll
  for (int i = 0; i < CPS; ++i) {
    const int c = c0 + i;
    if (act && c < c1) {
      const double Pc = c == C - 1 ? PLl : PL;
      Ex = ec[i].x + Pc * Ex;
      Ey = ec[i].y + Pc * Ey;
      PP *= Pc;
    }
  }
  segE[0][seg][el] = Ex; segE[1][seg][el] = Ey; segP[seg][el] = PP;
  __syncthreads();
  double ax = 0.0, ay = 0.0, tot = 1.0;
#pragma unroll
  for (int j = 0; j < kCarrySeg; ++j) {  // segments seg, seg+1, ..., seg-1: the value entering segment `seg`
    const int s2 = (seg + j) & (kCarrySeg - 1);
    ax = segE[0][s2][el] + segP[s2][el] * ax;
    ay = segE[1][s2][el] + segP[s2][el] * ay;
    tot *= segP[s2][el];
  }
  const double close = 1.0 / (1.0 - tot);  // tot = rho^nx < 1 for every k >= 1
  ax *= close; ay *= close;
  __syncthreads();
  // ---- forward walk with the incoming value; the backward segment map accumulates on the way
  //      (Bout[c0] = sum_i T_i prod_{i' < i} P_i' + prod_i P_i Bin)
  double Tx = 0.0, Ty = 0.0, W = 1.0;
#pragma unroll
  for (int i = 0; i < CPS; ++i) {
    const int c = c0 + i;
    if (act && c < c1) {
      const bool lastc = c == C - 1;
      const double Pc = lastc ? PLl : PL, Gc = lastc ? rG2l : rG2;
      const float2 ain = make_float2((float)ax, (float)ay);  // the float32 value K6 will use
      cA[(size_t)c * img + e] = ain;
      ax = ec[i].x + Pc * ax;
      ay = ec[i].y + Pc * ay;
      ec[i] = ain;
      Tx += (Sc[i].x + ain.x * Gc) * W;
      Ty += (Sc[i].y + ain.y * Gc) * W;
      W *= Pc;
    }
  }
  segE[0][seg][el] = Tx; segE[1][seg][el] = Ty; segP[seg][el] = W;
  __syncthreads();
  double bx = 0.0, by = 0.0;
#pragma unroll
  for (int j = 0; j < kCarrySeg; ++j) {  // segments seg, seg-1, ..., seg+1: the value entering `seg` from above
    const int s2 = (seg - j) & (kCarrySeg - 1);
    bx = segE[0][s2][el] + segP[s2][el] * bx;
    by = segE[1][s2][el] + segP[s2][el] * by;
  }
  bx *= close; by *= close;
#pragma unroll
  for (int i = CPS - 1; i >= 0; --i) {
    const int c = c0 + i;
    if (act && c < c1) {
      const bool lastc = c == C - 1;
      const double Pc = lastc ? PLl : PL, Gc = lastc ? rG2l : rG2;
      cB[(size_t)c * img + e] = make_float2((float)bx, (float)by);
      bx = (Sc[i].x + ec[i].x * Gc) + Pc * bx;
      by = (Sc[i].y + ec[i].y * Gc) + Pc * by;
    }
  }
}

// K6: backward sweep + inverse row transform + projection.  A CTA owns the chunk of K4 and walks it from its
// last row to its first in batches of NTB rows; the first batch also transforms the row AFTER the chunk (its
// spectrum is gain * cB: the u-correction of the chunk's last row needs q of that row).
template <int R1, int R2, int R3, int PS2, bool PROJECT>
__global__ void __launch_bounds__(1024, 1)
rows_inv_sweep_kernel(SweepParams P, const float2* __restrict__ tables, const float* u, const float* v, const float* p,
                      float* u_out, float* v_out, float* p_out) {
  using G = RowTeam<R1, R2, R3, PS2>;
  constexpr int N = G::N, T = G::T, SUB1 = G::SUB1;
  constexpr int NTM = G::NT < kSweepMaxTeams ? G::NT : kSweepMaxTeams;
  extern __shared__ __align__(16) float2 smr[];
  __shared__ __align__(8) unsigned long long tbar;
  float2* const ta = smr;
  float2* const tb = ta + G::TA;
  float2* const tu = tb + G::TB;
  float2* const bufs = smr + G::TAB;
  const int t = threadIdx.x, nthreads = blockDim.x, NTB = nthreads / T;
  const int team = t / T, b = t - team * T;
  float2* const buf = bufs + team * G::BUF;
  float2* const qsave = bufs + NTB * G::BUF;  // q of the row above the batch (the previous batch's lowest row)
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&tbar);
  pdl_trigger();
  if (t == 0) tables_issue<G>(smr, tables, bar);
  const GridF& g = P.g;
  const int ny = g.ny;
  const int i0 = blockIdx.x * P.L;
  const int Lc = min(P.L, g.nx - i0);
  const size_t plane = (size_t)g.nx * ny;
  if (PROJECT) { u += blockIdx.y * plane; v += blockIdx.y * plane; p += blockIdx.y * plane; u_out += blockIdx.y * plane; v_out += blockIdx.y * plane; }
  p_out += blockIdx.y * plane;
  const float2* const arow = P.a + ((size_t)blockIdx.y * g.nx + i0) * P.img;
  const float2* const cA = P.cA + ((size_t)blockIdx.y * P.C + blockIdx.x) * P.img;
  float2* const cB = P.cB + ((size_t)blockIdx.y * P.C + blockIdx.x) * P.img;
  pdl_wait();
  __syncthreads();  // barrier initialisation visible
  bool have_tables = false;
  // virtual rows j = 0 .. Lc of the chunk (j = Lc: the row after it), highest first
  for (int jhi = Lc; jhi >= 0; jhi -= NTB) {
    const int jbase = jhi - NTB + 1;       // row of team 0 (may be negative in the last batch)
    const int jlo = max(jbase, 0);
    const int j = jbase + team;            // this team's row
    const bool live = j >= 0;              // uniform over the team
    const bool more = jlo > 0;             // another batch follows
    // ---- backward sweep, elementwise over the image; q-spectra go to the team buffers
    for (int s = 0; s < kSweepSlots; ++s) {
      const int e = t + s * nthreads;
      if (e >= G::BUF) break;
      const float rho = __ldg(P.rho + e);
      if (rho < 0.0f) continue;
      const float lg = __ldg(P.lg + e), gain = __ldg(P.gain + e);
      const float2 ain = cA[e];
      float2 bn = cB[e];  // first batch: the carried value; later: b of the previous batch's lowest row
      float2* const col = bufs + e;
#pragma unroll
      for (int r = 0; r < NTM; ++r) {  // the loads are independent: all in flight before the first is used
        const int jr = jbase + r;
        if (r < NTB && jr >= 0 && jr < Lc) col[r * G::BUF] = arow[(size_t)jr * P.img + e];
      }
#pragma unroll
      for (int r = NTM - 1; r >= 0; --r) {
        const int jr = jbase + r;
        if (r < NTB && jr >= 0) {
          if (jr < Lc) {
            const float2 x = col[r * G::BUF];
            const float pw = exp2f((float)(jr + 1) * lg);  // rho^(jr+1): the carried part of the forward sweep
            bn.x = fmaf(rho, bn.x, fmaf(pw, ain.x, x.x));
            bn.y = fmaf(rho, bn.y, fmaf(pw, ain.y, x.y));
          }
          col[r * G::BUF] = make_float2(gain * bn.x, gain * bn.y);
        }
      }
      if (more) cB[e] = bn;
    }
    __syncthreads();
    if (live) {  // tangle IN PLACE: spectrum slots -> packed complex sequence at its digit-reversed positions
      if (!have_tables) { tables_wait(bar); have_tables = true; }
#pragma unroll
      for (int m = 0; m < (N / 2) / T; ++m) {
        const int k = b + T * m;
        if (k == 0) {
          const float x0 = buf[0].x, xn = buf[G::pad(N)].x;  // X[0], X[N]: both real
          buf[0] = make_float2(0.5f * (x0 + xn), 0.5f * (x0 - xn));
        } else {
          const int pk = G::pad(G::pos(k)), pm = G::pad(G::pos(N - k));
          const float2 a = buf[pk], c = cconj(buf[pm]);
          const float2 e = cscale(cadd(a, c), 0.5f);
          const float2 o = cmulc(cscale(csub(a, c), 0.5f), tu[k]);
          buf[pk] = cadd(e, mul_pi(o));
          buf[pm] = cadd(cconj(e), mul_pi(cconj(o)));
        }
      }
      if (b == 0) {
        const int pk = G::pad(G::pos(N / 2));
        const float2 a = buf[pk], c = cconj(a);
        const float2 e = cscale(cadd(a, c), 0.5f);
        const float2 o = cmulc(cscale(csub(a, c), 0.5f), tu[N / 2]);
        buf[pk] = cadd(e, mul_pi(o));
      }
    }
    team_bar<G>(team);
    team_inner<G, R2, R3, R1, true>(buf, tb, b, team, live);
    if (live) {
#pragma unroll
      for (int jj = 0; jj < G::NBA; ++jj) {  // pass A inverse; q[i][2n], q[i][2n+1] stay at (padded) index n
        const int bA = b + T * jj;
        float2* const e = buf + G::pad(bA);
        float2 z[R1];
#pragma unroll
        for (int r = 0; r < R1; ++r) {
          z[r] = e[G::pad(SUB1 * r)];
          if (r > 0) z[r] = cmulc(z[r], ta[(r - 1) * SUB1 + bA]);
        }
        dft_r<R1, true>(z);
        if (!PROJECT) {
          if (j < Lc) {
            float2* qo = reinterpret_cast<float2*>(p_out + (size_t)(i0 + j) * ny);
#pragma unroll
            for (int r = 0; r < R1; ++r) qo[bA + SUB1 * r] = z[r];
          }
        } else {
#pragma unroll
          for (int r = 0; r < R1; ++r) e[G::pad(SUB1 * r)] = z[r];
        }
      }
    }
    if (!PROJECT) {
      if (more) __syncthreads();
      continue;
    }
    __syncthreads();  // every team's q row is complete
    if (live && j < Lc) {
      const int ii = i0 + j;
      const float2* const q0row = buf;
      const float2* const q1row = team + 1 < NTB ? buf + G::BUF : qsave;  // q[i+1]
      const size_t ro = (size_t)ii * ny;
      const float2* u2 = reinterpret_cast<const float2*>(u + ro);
      const float2* v2 = reinterpret_cast<const float2*>(v + ro);
      const float2* p2 = reinterpret_cast<const float2*>(p + ro);
      float2* uo = reinterpret_cast<float2*>(u_out + ro);
      float2* vo = reinterpret_cast<float2*>(v_out + ro);
      float2* po = reinterpret_cast<float2*>(p_out + ro);
      constexpr int CH = 4;
#pragma unroll 1
      for (int m0 = 0; m0 < N / T; m0 += CH) {
        float2 ua[CH], va[CH], pa[CH];  // all loads first: the outputs may alias the inputs
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int n = b + T * (m0 + c);
          ua[c] = u2[n];
          va[c] = v2[n];
          pa[c] = p2[n];
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int n = b + T * (m0 + c);
          const float2 q0 = q0row[G::pad(n)], q1 = q1row[G::pad(n)];
          const float qn = q0row[G::pad(n + 1 < N ? n + 1 : 0)].x;  // q[i][2n+2] (periodic)
          float2 uu = ua[c], vv = va[c];
          uu.x -= (q1.x - q0.x) * g.inv_dx;
          uu.y -= (q1.y - q0.y) * g.inv_dx;
          vv.x -= (q0.y - q0.x) * g.inv_dy;
          vv.y -= (qn - q0.y) * g.inv_dy;
          uo[n] = uu;
          vo[n] = vv;
          po[n] = make_float2(q0.x + pa[c].x, q0.y + pa[c].y);
        }
      }
    }
    if (more) {  // keep q of the batch's lowest row for the next batch's highest one
      __syncthreads();
      if (team == 0)
        for (int n = b; n < G::BUF; n += T) qsave[n] = buf[n];
      __syncthreads();
    }
  }
}


// ---- x-sweep launchers
template <int R1, int R2, int R3, int PS2>
int launch_sweep_t(cudaStream_t s, const SweepParams& P, int ntb, const float2* tables, int stage, const float* u,
                   const float* v, const float* p, float* uo, float* vo, float* po, bool project) {
  using G = RowTeam<R1, R2, R3, PS2>;
  static bool configured = false;
  if (!configured) {
    const int big = (int)G::smem((G::NT < kSweepMaxTeams ? G::NT : kSweepMaxTeams) + 1);
    cudaFuncSetAttribute(rows_fwd_sweep_kernel<R1, R2, R3, PS2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(rows_inv_sweep_kernel<R1, R2, R3, PS2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(rows_inv_sweep_kernel<R1, R2, R3, PS2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    configured = true;
  }
  if (P.img != G::BUF) { set_error("x-sweep: row image %d does not match the kernel's %d", P.img, G::BUF); return JAXIB_ERR_INVALID; }
  dim3 grid(P.C, P.g.batch), block(ntb * G::T);
  if (stage == 0) {
    launch_pdl(8, rows_fwd_sweep_kernel<R1, R2, R3, PS2>, grid, block, G::smem(ntb), s, P, tables, u, v);
    return check_launch("rows_fwd_sweep_kernel");
  }
  if (project)
    launch_pdl(32, rows_inv_sweep_kernel<R1, R2, R3, PS2, true>, grid, block, G::smem(ntb + 1), s, P, tables, u, v, p, uo, vo, po);
  else
    launch_pdl(32, rows_inv_sweep_kernel<R1, R2, R3, PS2, false>, grid, block, G::smem(ntb + 1), s, P, tables, u, v, p, uo, vo, po);
  return check_launch("rows_inv_sweep_kernel");
}

int launch_sweep_stage(cudaStream_t s, const SweepParams& P, int ntb, const float2* tables, int stage, const float* u,
                       const float* v, const float* p, float* uo, float* vo, float* po, bool project) {
  switch (P.g.ny / 2) {
    case 512: return launch_sweep_t<8, 8, 8, 7>(s, P, ntb, tables, stage, u, v, p, uo, vo, po, project);
    case 1024: return launch_sweep_t<8, 16, 8, 8>(s, P, ntb, tables, stage, u, v, p, uo, vo, po, project);
    case 2048: return launch_sweep_t<16, 16, 8, 8>(s, P, ntb, tables, stage, u, v, p, uo, vo, po, project);
    case 4096: return launch_sweep_t<16, 16, 16, 8>(s, P, ntb, tables, stage, u, v, p, uo, vo, po, project);
    default: set_error("x-sweep: unsupported row length %d", P.g.ny); return JAXIB_ERR_UNSUPPORTED;
  }
}

}  // namespace

// radix triples of the row team kernels for a complex length n = ny/2 (0: unsupported)
bool team_rows_supported(int n) { return n == 512 || n == 1024 || n == 2048 || n == 4096; }
void team_row_radices(int n, int* r) {
  switch (n) {
    case 512: r[0] = 8; r[1] = 8; r[2] = 8; break;
    case 1024: r[0] = 8; r[1] = 16; r[2] = 8; break;
    case 2048: r[0] = 16; r[1] = 16; r[2] = 8; break;
    case 4096: r[0] = 16; r[1] = 16; r[2] = 16; break;
    default: r[0] = r[1] = r[2] = 0;
  }
  r[3] = 0;
}

// [TA | TB | TU] in the kernels' shared-memory layout; the row length is ny = 2 n
void team_row_twiddle_table(int n, std::vector<float2>* out) {
  int r[4];
  team_row_radices(n, r);
  const int R1 = r[0], R2 = r[1];
  const int sub1 = n / R1, sub2 = sub1 / R2;
  const double two_pi = 6.283185307179586476925286766559;
  auto w = [&](long long m, int len) {
    m %= len;
    return make_float2((float)cos(two_pi * m / len), (float)-sin(two_pi * m / len));
  };
  out->clear();
  for (int q = 1; q < R1; ++q)
    for (int b = 0; b < sub1; ++b) out->push_back(w((long long)b * q, n));
  for (int q = 1; q < R2; ++q)
    for (int b2 = 0; b2 < sub2; ++b2) out->push_back(w((long long)b2 * q * R1, n));
  for (int k = 0; k <= n / 2; ++k) out->push_back(w(k, 2 * n));
  while (out->size() % 2) out->push_back(make_float2(0.f, 0.f));
}

int launch_rows_fwd_team(cudaStream_t s, const FastRowParams& P, const float2* tables, const float* u, const float* v,
                         float2* spec) {
  switch (P.g.ny / 2) {
    case 512: return launch_rows_fwd_team_t<8, 8, 8, 7>(s, P, tables, u, v, spec);
    case 1024: return launch_rows_fwd_team_t<8, 16, 8, 8>(s, P, tables, u, v, spec);
    case 2048: return launch_rows_fwd_team_t<16, 16, 8, 8>(s, P, tables, u, v, spec);
    case 4096: return launch_rows_fwd_team_t<16, 16, 16, 8>(s, P, tables, u, v, spec);
    default: set_error("team row kernels: unsupported row length %d", P.g.ny); return JAXIB_ERR_UNSUPPORTED;
  }
}

int launch_rows_inv_team(cudaStream_t s, const FastRowParams& P, const float2* tables, const float2* spec, const float* u,
                         const float* v, const float* p, float* uo, float* vo, float* po, bool project) {
  switch (P.g.ny / 2) {
    case 512: return launch_rows_inv_team_t<8, 8, 8, 7>(s, P, tables, spec, u, v, p, uo, vo, po, project);
    case 1024: return launch_rows_inv_team_t<8, 16, 8, 8>(s, P, tables, spec, u, v, p, uo, vo, po, project);
    case 2048: return launch_rows_inv_team_t<16, 16, 8, 8>(s, P, tables, spec, u, v, p, uo, vo, po, project);
    case 4096: return launch_rows_inv_team_t<16, 16, 16, 8>(s, P, tables, spec, u, v, p, uo, vo, po, project);
    default: set_error("team row kernels: unsupported row length %d", P.g.ny); return JAXIB_ERR_UNSUPPORTED;
  }
}


// ---- x-sweep: geometry, tables, launchers (see the kernels above)
template <int R1, int R2, int R3, int PS2>
static void sweep_layout_t(SweepGeom* out, std::vector<int>* slot_of_k) {
  using G = RowTeam<R1, R2, R3, PS2>;
  out->img = G::BUF;
  out->team_threads = G::T;
  out->max_teams = G::NT < kSweepMaxTeams ? G::NT : kSweepMaxTeams;
  slot_of_k->resize(G::N + 1);
  for (int k = 0; k < G::N; ++k)
    (*slot_of_k)[k] = G::pad((k % R1) * G::SUB1 + ((k / R1) % R2) * G::SUB2 + k / (R1 * R2));
  (*slot_of_k)[G::N] = G::pad(G::N);
}

static bool sweep_layout(int n2, SweepGeom* out, std::vector<int>* slot_of_k) {
  switch (n2) {
    case 512: sweep_layout_t<8, 8, 8, 7>(out, slot_of_k); return true;
    case 1024: sweep_layout_t<8, 16, 8, 8>(out, slot_of_k); return true;
    case 2048: sweep_layout_t<16, 16, 8, 8>(out, slot_of_k); return true;
    case 4096: sweep_layout_t<16, 16, 16, 8>(out, slot_of_k); return true;
    default: return false;
  }
}

// Chunk length and teams per CTA.  One CTA per chunk; a CTA transforms `ntb` rows at a time (K6: ntb - 1 rows of
// the chunk plus the row after it in its first batch).  Small grids: fewer teams per CTA until every SM has a
// chunk.  Large single grids: one wave of CTAs, each walking its chunk in several batches (the carried values
// cost 3 / L of the spectrum traffic, so chunks should not be short either).
bool sweep_geometry(int nx, int ny, int batch, SweepGeom* out) {
  std::vector<int> slots;
  if (ny % 4 || !sweep_layout(ny / 2, out, &slots) || nx < 2) return false;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long rows = (long long)nx * batch;
  int ntb = out->max_teams;
  while (ntb > 4 && (rows + ntb - 2) / (ntb - 1) < (3 * sms) / 4) ntb /= 2;  // fewer than 3/4 of the SMs would get a chunk
  if (const char* e = getenv("JAXIB_SWEEP_NTB")) { const int v = atoi(e); if (v >= 4 && v <= out->max_teams) ntb = v; }
  int L = ntb - 1;
  if (batch == 1 && (nx + L - 1) / L > sms) L = (nx + sms - 1) / sms;
  if (const char* e = getenv("JAXIB_SWEEP_L")) { const int v = atoi(e); if (v >= 1) L = v; }
  if ((nx + L - 1) / L > 36 * 16) L = (nx + 36 * 16 - 1) / (36 * 16);  // the carry kernel walks at most 36 chunks per thread
  if (L > nx) L = nx;
  out->ntb = ntb;
  out->L = L;
  out->C = (nx + L - 1) / L;
  return true;
}

void sweep_tables(int ny, double dx, double dy, std::vector<float>* rho, std::vector<float>* lg, std::vector<float>* gain,
                  std::vector<double>* rho_d) {
  SweepGeom geo;
  std::vector<int> slots;
  const int n2 = ny / 2;
  sweep_layout(n2, &geo, &slots);
  rho->assign(geo.img, -1.0f);
  lg->assign(geo.img, 0.0f);
  gain->assign(geo.img, 0.0f);
  rho_d->assign(geo.img, -1.0);
  const double two_pi = 6.283185307179586476925286766559;
  for (int k = 0; k <= n2; ++k) {
    const int e = slots[k];
    if (k == 0) {  // the carry kernel solves this column; the row kernels pass it through
      (*rho)[e] = 0.0f; (*rho_d)[e] = 0.0; (*lg)[e] = -1.0e30f; (*gain)[e] = (float)(1.0 / n2);
      continue;
    }
    const double mu = (2.0 - 2.0 * cos(two_pi * k / ny)) * (dx * dx) / (dy * dy);
    const double r = 1.0 / (1.0 + 0.5 * mu + sqrt(mu + 0.25 * mu * mu));  // the root inside the unit circle
    (*rho_d)[e] = r;
    (*rho)[e] = (float)r;
    (*lg)[e] = (float)log2(r);
    (*gain)[e] = (float)(-dx * dx * r / n2);
  }
}

int launch_rows_fwd_sweep(cudaStream_t s, const SweepParams& P, int ntb, const float2* tables, const float* u, const float* v) {
  return launch_sweep_stage(s, P, ntb, tables, 0, u, v, nullptr, nullptr, nullptr, nullptr, false);
}

int launch_rows_inv_sweep(cudaStream_t s, const SweepParams& P, int ntb, const float2* tables, const float* u, const float* v,
                          const float* p, float* uo, float* vo, float* po, bool project) {
  return launch_sweep_stage(s, P, ntb, tables, 1, u, v, p, uo, vo, po, project);
}

int launch_sweep_carry(cudaStream_t s, const SweepParams& P) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(sweep_carry_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(sweep_carry_kernel<36>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured = true;
  }
  const size_t smem = (size_t)P.g.nx * sizeof(double);
  if (smem > 200 * 1024) { set_error("x-sweep: nx=%d exceeds the k = 0 column buffer", P.g.nx); return JAXIB_ERR_UNSUPPORTED; }
  const int cps = (P.C + kCarrySeg - 1) / kCarrySeg;
  if (cps > 36) { set_error("x-sweep: %d chunks exceed the carry kernel's %d", P.C, 36 * kCarrySeg); return JAXIB_ERR_UNSUPPORTED; }
  dim3 grid((P.img + 31) / 32 + 1, P.g.batch);
  if (cps <= 10) launch_pdl(16, sweep_carry_kernel<10>, grid, dim3(kCarryThreads), smem, s, P);
  else launch_pdl(16, sweep_carry_kernel<36>, grid, dim3(kCarryThreads), smem, s, P);
  return check_launch("sweep_carry_kernel");
}

}  // namespace jaxib
