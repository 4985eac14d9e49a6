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
  dim3 grid((P.g.nx + rows - 1) / rows, P.g.batch);
  if (project)
    launch_pdl(32, rows_inv_team_kernel<R1, R2, R3, PS2, true>, grid, dim3(nt * G::T), G::smem(nt), s, P, tables, spec, u, v, p, uo, vo, po);
  else
    launch_pdl(32, rows_inv_team_kernel<R1, R2, R3, PS2, false>, grid, dim3(nt * G::T), G::smem(nt), s, P, tables, spec, u, v, p, uo, vo, po);
  return check_launch("rows_inv_team_kernel");
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

}  // namespace jaxib
