// Power-of-two fast paths of the Poisson kernels (N = R1*R2*R3 with radices 8/16, i.e. 512..4096).
//
// Same mathematics as poisson.cu (forward DIF, diagonal in digit-reversed positions, exact inverse),
// re-organised so that the FIRST forward pass reads straight from global memory into registers and
// the LAST inverse pass writes straight from registers to global memory; only the two inner passes
// go through shared memory, with compile-time strides (no div/mod) and the innermost pass fused
// with the diagonal multiply:   [pass3 fwd] * D * [pass3 inv]  in registers.
//
//   cols_fast  (K5): a CTA owns 4 adjacent ky columns = two float4 "column pairs"; every global
//                    access is a 16-byte load/store and each 32-byte sector of the [nx][sp] spectrum
//                    is used completely by the two threads that share it.
//   rows_fwd_fast (K4) / rows_inv_fast (K6): one 64..256-thread group per row, several rows per CTA
//                    in sequence; the untangle of the packed real transform reads the digit-reversed
//                    positions with shifts and masks.
#include <stdlib.h>

#include "fft.cuh"
#include "poisson_fast.cuh"

namespace jaxib {

namespace {

constexpr float kCutoffF = 10.0f * 1.1920928955078125e-07f;

__device__ __forceinline__ int padq(int p) { return p + (p >> 3); }  // float4 (two-column) arrays

template <int R, bool INV>
__device__ __forceinline__ void dft_dir(float2* v) {
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

// ---- cp.async (LDGSTS) staging: global -> shared without passing through registers ---------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(PENDING));
}
// cooperative copy of `bytes` (multiple of 16, both sides 16-byte aligned) by `nthreads` threads
__device__ __forceinline__ void copy_async(void* dst, const void* src, int bytes, int tid, int nthreads) {
  char* d = static_cast<char*>(dst);
  const char* s = static_cast<const char*>(src);
  for (int o = tid * 16; o < bytes; o += nthreads * 16) cp_async16(d + o, s + o);
}

__device__ __forceinline__ float inv_eig_f(float l) { return fabsf(l) > kCutoffF ? __frcp_rn(l) : 0.0f; }

// ---------------------------------------------------------------------------------------------- K5
// float4 = two adjacent ky columns at one kx.  Bank swizzle of the [N] float4 column-pair buffer:
// XOR index bits 4-6 into bits 0-2 so that the stride-R3 accesses of the innermost pass spread over the
// eight 16-byte bank groups (contiguous runs stay contiguous inside aligned blocks of 16).
__device__ __forceinline__ int qswz(int p) { return p ^ ((p >> 4) & 7); }

// Thread mapping: the global (first / last) pass moves float4 = both columns of a pair with S1 = N/R1
// threads per pair (R1 = 8 keeps it light in registers); the two inner passes run ONE column per
// thread (8-byte shared accesses), so a radix-16 butterfly costs 32 data registers instead of 64 and
// every thread has work in every pass.
template <int R1, int R2, int R3, int NP>
__global__ void __launch_bounds__(NP * R2 * R3, (NP * R2 * R3 <= 256 ? 2 : 1))
cols_fast_kernel(FastColParams P, float2* __restrict__ spec) {
  constexpr int N = R1 * R2 * R3;
  constexpr int S1 = N / R1;   // threads per column pair; stride of pass 1
  constexpr int L2 = R2 * R3;  // block length of pass 2
  extern __shared__ float4 smq[];
  constexpr int npairs = NP;  // column pairs per CTA: 4 pairs = 64 contiguous bytes per spectrum row
  // inner passes: thread -> (pair, b) blocked, so that a warp stays inside one pair's buffer
  const int pair = threadIdx.x / S1, b = threadIdx.x - pair * S1;
  const int c = blockIdx.x * 2 * npairs + 2 * pair;  // columns c, c+1 (padding columns hold zeros)
  // pair buffers are N + 2 float4 apart: with the pair index fastest across lanes a quarter-warp's 16-byte
  // accesses (4 pairs x 2 rows) then hit 8 different bank groups instead of 2 (4-way conflict measured)
  constexpr int PQ = N + 2;
  float4* col = smq + pair * PQ;
  float2* col2 = reinterpret_cast<float2*>(col);
  // global passes: thread -> (b, pair) with the pair index FASTEST, so that adjacent lanes cover the
  // adjacent 16-byte halves of one 32-byte sector of a spectrum row (a warp-wide load then fetches
  // every sector once; with the blocked mapping each sector was fetched by two different warps)
  const int gpair = threadIdx.x % npairs, gb = threadIdx.x / npairs;
  float4* gcol = smq + gpair * PQ;
  const int gc = blockIdx.x * 2 * npairs + 2 * gpair;  // first column of this thread's global pair
  const bool gvalid = gc + 1 < P.limit;                // the last CTA may overhang the columns of the buffer
  float4* g4 = reinterpret_cast<float4*>(spec + (size_t)blockIdx.y * P.nx * P.sp + gc);
  const size_t rs = (size_t)P.sp / 2;  // row stride in float4 units
  const float2* __restrict__ tw = P.tw;
  // twiddle tables in shared memory (coalesced fill): W_N^m for the outer pass, W_N^(b2 r R1) for pass 2.
  // Reading them from global with a lane stride of r*8 bytes costs up to 32 L1 wavefronts per load.
  float2* const tws = reinterpret_cast<float2*>(smq + npairs * PQ);
  float2* const tw2 = tws + N;

  {  // pass 1 forward: global -> registers -> shared
    // prologue (independent of the previous kernel): the twiddle tables; then wait for the spectrum
    // CTAs of the first wave may start under the tail of the previous kernel: they fill the tables before
    // waiting for it; later waves fill them while their column loads are in flight
    pdl_trigger();
    const bool early = (int)blockIdx.x < P.first_wave;
    auto fill_tables = [&]() {
      for (int e = threadIdx.x; e < N; e += blockDim.x) tws[e] = __ldg(tw + e);
      for (int e = threadIdx.x; e < R3 * R2; e += blockDim.x) tw2[(e / R2) * (R2 + 1) + e % R2] = __ldg(tw + (e / R2) * (e % R2) * R1);
    };
    if (early) fill_tables();
    pdl_wait();
    float2 va[R1], vb[R1];
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gvalid) x = g4[(size_t)(gb + r * S1) * rs];
      va[r] = make_float2(x.x, x.y);
      vb[r] = make_float2(x.z, x.w);
    }
    if (!early) fill_tables();
    __syncthreads();
    Dft<R1>::run(va);
    Dft<R1>::run(vb);
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      if (r > 0) {
        const float2 w = tws[gb * r];
        va[r] = cmul(va[r], w);
        vb[r] = cmul(vb[r], w);
      }
      gcol[qswz(gb + r * S1)] = make_float4(va[r].x, va[r].y, vb[r].x, vb[r].y);
    }
  }
  __syncthreads();
  {
  // pass 2 forward, one column per thread
#pragma unroll
  for (int m = 0; m < (2 * R1 + R2 - 1) / R2; ++m) {
    const int w = b + m * S1;
    if (w >= 2 * (N / R2)) continue;
    const int cc = w & 1, bf = w >> 1;
    const int blk = bf / R3, b2 = bf - blk * R3;
    const int p0 = blk * L2 + b2;
    float2 v[R2];
#pragma unroll
    for (int r = 0; r < R2; ++r) v[r] = col2[2 * qswz(p0 + r * R3) + cc];
    Dft<R2>::run(v);
#pragma unroll
    for (int r = 1; r < R2; ++r) v[r] = cmul(v[r], tw2[b2 * (R2 + 1) + r]);
#pragma unroll
    for (int r = 0; r < R2; ++r) col2[2 * qswz(p0 + r * R3) + cc] = v[r];
  }
  __syncthreads();
  // pass 3 forward, diagonal, pass 3 inverse -- in registers, one column per thread
#pragma unroll
  for (int m = 0; m < (2 * R1 + R3 - 1) / R3; ++m) {
    const int w = b + m * S1;
    if (w >= 2 * (N / R3)) continue;
    const int cc = w & 1, p0 = (w >> 1) * R3;
    const float ly = P.lam_y[c + cc];
    float2 v[R3];
#pragma unroll
    for (int r = 0; r < R3; ++r) v[r] = col2[2 * qswz(p0 + r) + cc];
    Dft<R3>::run(v);
    // eigenvalues are fetched four at a time right where they are used (L1-resident table): holding all
    // R3 of them across the butterfly cost 16 registers and pushed the 64-register kernel into spills
    const float4* lx4 = reinterpret_cast<const float4*>(P.lam_x_pos + p0);
#pragma unroll
    for (int r = 0; r < R3; r += 4) {
      const float4 l4 = __ldg(lx4 + r / 4);
      v[r] = cscale(v[r], inv_eig_f(l4.x + ly));
      v[r + 1] = cscale(v[r + 1], inv_eig_f(l4.y + ly));
      v[r + 2] = cscale(v[r + 2], inv_eig_f(l4.z + ly));
      v[r + 3] = cscale(v[r + 3], inv_eig_f(l4.w + ly));
    }
    dft_dir<R3, true>(v);
#pragma unroll
    for (int r = 0; r < R3; ++r) col2[2 * qswz(p0 + r) + cc] = v[r];
  }
  __syncthreads();
  // pass 2 inverse
#pragma unroll
  for (int m = 0; m < (2 * R1 + R2 - 1) / R2; ++m) {
    const int w = b + m * S1;
    if (w >= 2 * (N / R2)) continue;
    const int cc = w & 1, bf = w >> 1;
    const int blk = bf / R3, b2 = bf - blk * R3;
    const int p0 = blk * L2 + b2;
    float2 v[R2];
#pragma unroll
    for (int r = 0; r < R2; ++r) v[r] = col2[2 * qswz(p0 + r * R3) + cc];
#pragma unroll
    for (int r = 1; r < R2; ++r) v[r] = cmulc(v[r], tw2[b2 * (R2 + 1) + r]);
    dft_dir<R2, true>(v);
#pragma unroll
    for (int r = 0; r < R2; ++r) col2[2 * qswz(p0 + r * R3) + cc] = v[r];
  }
  }
  __syncthreads();
  {  // pass 1 inverse: shared -> registers -> global
    float2 va[R1], vb[R1];
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      float4 x = gcol[qswz(gb + r * S1)];
      va[r] = make_float2(x.x, x.y);
      vb[r] = make_float2(x.z, x.w);
      if (r > 0) {
        const float2 w = tws[gb * r];
        va[r] = cmulc(va[r], w);
        vb[r] = cmulc(vb[r], w);
      }
    }
    dft_dir<R1, true>(va);
    dft_dir<R1, true>(vb);
    if (gvalid) {
#pragma unroll
      for (int r = 0; r < R1; ++r)
        g4[(size_t)(gb + r * S1) * rs] = make_float4(va[r].x, va[r].y, vb[r].x, vb[r].y);
    }
  }
}

// ---------------------------------------------------------------------------------------------- rows
// Bank swizzle of a row buffer (8-byte elements: the low 4 index bits select the bank pair).  Besides
// contiguous runs the row kernels access it with stride R3 (innermost pass), stride L2 = R2*R3 in the
// fastest lanes (the digit-reversed scatter/gather of the real-FFT untangle) and b2 + L2*blk (pass 2);
// XOR-ing index bits 4-6 and 7-9 into bits 0-2 and bit 7 into bit 3 spreads all of them over the 16
// bank pairs (plain padding left the untangle 8-way conflicted: half of all shared wavefronts).
__device__ __forceinline__ int rswz(int p) {
  return p ^ (((p >> 4) ^ (p >> 7)) & 7) ^ (((p >> 7) & 1) << 3);
}

template <int R1, int R2, int R3>
struct RowGeom {
  static constexpr int N = R1 * R2 * R3;  // complex length ny/2
  static constexpr int S1 = N / R1;       // threads per row group
  static constexpr int L2 = R2 * R3;
  static constexpr int PITCH = N + 16;  // a multiple of 128 bytes: the cp.async staging buffers behind the rows stay line-aligned
  // position of frequency k after the forward DIF (digit reversal of the radix triple)
  __device__ static __forceinline__ int pos(int k) {
    return (k % R1) * L2 + ((k / R1) % R2) * R3 + k / (R1 * R2);
  }
};

// inner passes (2 and 3) of the row transform on one shared-memory row; G = threads of the group
template <int R1, int R2, int R3, bool INV>
__device__ __forceinline__ void row_inner_passes(float2* row, const float2* tw2, int b) {
  using Gm = RowGeom<R1, R2, R3>;
  // N/R2 (N/R3) butterflies are shared by the S1 = N/R1 threads: each thread runs ceil(R1/R2) of them
  // (when R1 < R2 only the first N/R2 threads have one)
  constexpr int TRIPS2 = (R1 + R2 - 1) / R2, TRIPS3 = (R1 + R3 - 1) / R3;
  auto pass2 = [&]() {
#pragma unroll
    for (int m = 0; m < TRIPS2; ++m) {
      const int t = b + m * Gm::S1;
      if (t >= Gm::N / R2) continue;
      const int blk = t / R3, b2 = t - blk * R3;
      const int p0 = blk * Gm::L2 + b2;
      float2 v[R2];
#pragma unroll
      for (int r = 0; r < R2; ++r) v[r] = row[rswz(p0 + r * R3)];
      if (!INV) {
        Dft<R2>::run(v);
#pragma unroll
        for (int r = 1; r < R2; ++r) v[r] = cmul(v[r], tw2[b2 * (R2 + 1) + r]);
      } else {
#pragma unroll
        for (int r = 1; r < R2; ++r) v[r] = cmulc(v[r], tw2[b2 * (R2 + 1) + r]);
        dft_dir<R2, true>(v);
      }
#pragma unroll
      for (int r = 0; r < R2; ++r) row[rswz(p0 + r * R3)] = v[r];
    }
  };
  auto pass3 = [&]() {
#pragma unroll
    for (int m = 0; m < TRIPS3; ++m) {
      if (b + m * Gm::S1 >= Gm::N / R3) continue;
      const int p0 = (b + m * Gm::S1) * R3;
      float2 v[R3];
#pragma unroll
      for (int r = 0; r < R3; ++r) v[r] = row[rswz(p0 + r)];
      dft_dir<R3, INV>(v);
#pragma unroll
      for (int r = 0; r < R3; ++r) row[rswz(p0 + r)] = v[r];
    }
  };
  if (!INV) {
    pass2();
    __syncthreads();
    pass3();
    __syncthreads();
  } else {
    pass3();
    __syncthreads();
    pass2();
    __syncthreads();
  }
}

// K4: divergence + forward packed-real FFT of one row per iteration; the CTA is ONE row group.
template <int R1, int R2, int R3>
__global__ void __launch_bounds__(R2 * R3) rows_fwd_fast_kernel(FastRowParams P, const float* __restrict__ u,
                                                                const float* __restrict__ v,
                                                                float2* __restrict__ spec) {
  using Gm = RowGeom<R1, R2, R3>;
  constexpr int N = Gm::N, S1 = Gm::S1;
  // shared memory: FFT row (Gm::PITCH float2) | u rows i-1 / i (2 x ny floats, rotating) | v row i (ny floats)
  extern __shared__ float2 row[];
  pdl_trigger();
  const GridF& g = P.g;
  const int ny = g.ny;
  float* const stage = reinterpret_cast<float*>(row + Gm::PITCH);
  float* const ubuf[2] = {stage, stage + ny};
  float* const vbuf = stage + 2 * ny;
  const size_t plane = (size_t)g.nx * ny;
  u += blockIdx.y * plane;
  v += blockIdx.y * plane;
  spec += (size_t)blockIdx.y * g.nx * P.sp;
  const int b = threadIdx.x;
  const int i0 = blockIdx.x * P.rows_per_group;
  const int i1 = min(i0 + P.rows_per_group, g.nx);
  const float2* __restrict__ tw = P.tw;
  const int row_bytes = ny * (int)sizeof(float);
  // twiddles are row-independent: pass-1 factors W^(b r) live in registers, the small pass-2 table
  // W^(b2 r R1) in shared memory.  (Fetching them per row through L1 with a lane stride of r*8 bytes
  // costs up to 32 wavefronts per load and saturated the L1tex pipe.)
  float2* const tw2 = reinterpret_cast<float2*>(stage + 3 * ny);
  for (int e = b; e < R3 * R2; e += S1) tw2[(e / R2) * (R2 + 1) + e % R2] = __ldg(tw + 2 * (e / R2) * (e % R2) * R1);
  float2 tw1[R1];
#pragma unroll
  for (int r = 1; r < R1; ++r) tw1[r] = __ldg(tw + 2 * b * r);
  pdl_wait();  // everything above is independent of the previous kernel (twiddle prologue)
  // cp.async pipeline: the rows of step i+1 stream into shared memory while row i is transformed
  // slab mode (nx_global > 0): the rows are open -- u row -1 / spectrum row nx sit next to the arrays (halo)
  const ptrdiff_t im1 = (i0 == 0 && g.nx_global == 0) ? g.nx - 1 : i0 - 1;
  copy_async(ubuf[0], u + im1 * ny, row_bytes, b, S1);
  copy_async(ubuf[1], u + (size_t)i0 * ny, row_bytes, b, S1);
  copy_async(vbuf, v + (size_t)i0 * ny, row_bytes, b, S1);
  cp_async_commit();
  for (int i = i0; i < i1; ++i) {
    const int kk = i - i0;
    const float2* ur = reinterpret_cast<const float2*>(ubuf[(kk + 1) & 1]);
    const float2* um = reinterpret_cast<const float2*>(ubuf[kk & 1]);
    const float2* vr = reinterpret_cast<const float2*>(vbuf);
    cp_async_wait<0>();
    __syncthreads();
    float2 z[R1];
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      const int n = b + r * S1;
      const float2 a = ur[n], c = um[n], d = vr[n];
      const float vm = vbuf[n > 0 ? 2 * n - 1 : ny - 1];
      z[r] = make_float2((a.x - c.x) * g.inv_dx + (d.x - vm) * g.inv_dy, (a.y - c.y) * g.inv_dx + (d.y - d.x) * g.inv_dy);
    }
    __syncthreads();  // everyone has consumed u[i-1] and v[i]: their slots can be refilled
    if (i + 1 < i1) {
      copy_async(ubuf[kk & 1], u + (size_t)(i + 1) * ny, row_bytes, b, S1);
      copy_async(vbuf, v + (size_t)(i + 1) * ny, row_bytes, b, S1);
    }
    cp_async_commit();
    Dft<R1>::run(z);
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      if (r > 0) z[r] = cmul(z[r], tw1[r]);
      row[rswz(b + r * S1)] = z[r];
    }
    __syncthreads();
    row_inner_passes<R1, R2, R3, false>(row, tw2, b);
    // untangle: X[k] = E + W^k O, X[N-k] = conj(E - W^k O)
    float2* out = spec + (size_t)i * P.sp;
    const bool scatter = P.sc.cpr != 0;  // slab mode: column k lives in the buffer of the slab that owns it
    auto put = [&](int k, float2 val) {
      if (scatter) *spec_addr(P.sc, i, k) = val;
      else out[k] = val;
    };
#pragma unroll
    for (int m = 0; m <= (N / 2) / S1; ++m) {  // k = 0 .. N/2 (the last trip only carries k = N/2)
      const int k = b + m * S1;
      if (k > N / 2) continue;
      if (k == 0) {
        const float2 zz = row[rswz(0)];
        put(0, make_float2(zz.x + zz.y, 0.0f));
        put(N, make_float2(zz.x - zz.y, 0.0f));
      } else {
        const int km = N - k;
        const float2 zk = row[rswz(Gm::pos(k))], zm = cconj(row[rswz(Gm::pos(km))]);
        const float2 e = cscale(cadd(zk, zm), 0.5f);
        const float2 o = cscale(mul_mi(csub(zk, zm)), 0.5f);
        const float2 t = cmul(__ldg(tw + k), o);
        put(k, cadd(e, t));
        if (km != k) put(km, cconj(csub(e, t)));
      }
    }
    if (b < 3) put(N + 1 + b, make_float2(0.0f, 0.0f));  // padding columns feed cols_fast: keep them finite
    __syncthreads();
  }
}

// K6: inverse packed-real FFT fused with gradient, velocity correction and pressure update.
template <int R1, int R2, int R3, bool PROJECT>
__global__ void __launch_bounds__(R2 * R3, (R2 * R3 <= 128 ? 4 : 1)) rows_inv_fast_kernel(FastRowParams P, const float2* __restrict__ spec,
                                                                const float* u, const float* v, const float* p,
                                                                float* u_out, float* v_out, float* p_out) {
  using Gm = RowGeom<R1, R2, R3>;
  constexpr int N = Gm::N, S1 = Gm::S1;
  // shared memory (float2 units): two FFT rows | staged spectrum row (N + 4)
  extern __shared__ float2 rows_sm[];
  pdl_trigger();
  float2* const spec_s = rows_sm + 2 * Gm::PITCH;  // 128-byte aligned
  const GridF& g = P.g;
  const int ny = g.ny;
  const size_t plane = (size_t)g.nx * ny;
  spec += (size_t)blockIdx.y * g.nx * P.sp;
  if (PROJECT) { u += blockIdx.y * plane; v += blockIdx.y * plane; p += blockIdx.y * plane; u_out += blockIdx.y * plane; v_out += blockIdx.y * plane; }
  p_out += blockIdx.y * plane;
  const int b = threadIdx.x;
  const int i0 = blockIdx.x * P.rows_per_group;
  const int i1 = min(i0 + P.rows_per_group, g.nx);
  const int last = PROJECT ? i1 : i1 - 1;
  const float2* __restrict__ tw = P.tw;
  const int spec_bytes = (N + 4) * (int)sizeof(float2);
  float2* const tw2 = spec_s + N + 4;  // pass-2 twiddle table, see rows_fwd_fast_kernel
  for (int e = b; e < R3 * R2; e += S1) tw2[(e / R2) * (R2 + 1) + e % R2] = __ldg(tw + 2 * (e / R2) * (e % R2) * R1);
  float2 tw1[R1];
#pragma unroll
  for (int r = 1; r < R1; ++r) tw1[r] = __ldg(tw + 2 * b * r);
  float2 zprev[R1];
  pdl_wait();  // twiddle prologue done; the spectrum is the previous kernel's output
  // cp.async pipeline: the next spectrum row streams into shared memory behind the FFT passes
  copy_async(spec_s, spec + (size_t)i0 * P.sp, spec_bytes, b, S1);
  cp_async_commit();
  for (int ii = i0; ii <= last; ++ii) {
    const int i = (ii >= g.nx && g.nx_global == 0) ? ii - g.nx : ii;
    float2* cur = rows_sm + ((ii - i0) & 1) * Gm::PITCH;
    cp_async_wait<0>();  // the spectrum row issued one iteration ago has landed
    __syncthreads();
    const float2* in = spec_s;
#pragma unroll
    for (int m = 0; m <= (N / 2) / S1; ++m) {
      const int k = b + m * S1;
      if (k > N / 2) continue;
      if (k == 0) {
        const float x0 = in[0].x, xn = in[N].x;
        cur[rswz(0)] = make_float2(0.5f * (x0 + xn), 0.5f * (x0 - xn));
      } else {
        const int km = N - k;
        const float2 xk = in[k], xm = cconj(in[km]);
        const float2 e = cscale(cadd(xk, xm), 0.5f);
        const float2 o = cmulc(cscale(csub(xk, xm), 0.5f), __ldg(tw + k));
        cur[rswz(Gm::pos(k))] = cadd(e, mul_pi(o));
        if (km != k) cur[rswz(Gm::pos(km))] = cadd(cconj(e), mul_pi(cconj(o)));
      }
    }
    __syncthreads();
    if (ii < last) {  // the staged spectrum row is consumed: prefetch the next one behind the FFT passes
      const int inext = (ii + 1 >= g.nx && g.nx_global == 0) ? ii + 1 - g.nx : ii + 1;
      copy_async(spec_s, spec + (size_t)inext * P.sp, spec_bytes, b, S1);
    }
    cp_async_commit();
    row_inner_passes<R1, R2, R3, true>(cur, tw2, b);
    float2 z[R1];
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      z[r] = cur[rswz(b + r * S1)];
      if (r > 0) z[r] = cmulc(z[r], tw1[r]);
    }
    dft_dir<R1, true>(z);
#pragma unroll
    for (int r = 0; r < R1; ++r) z[r] = cscale(z[r], P.scale);  // q[i][2n], q[i][2n+1], n = b + r*S1
    if (!PROJECT) {
      float2* qo = reinterpret_cast<float2*>(p_out + (size_t)i * ny);
#pragma unroll
      for (int r = 0; r < R1; ++r) qo[b + r * S1] = z[r];
    } else {
      // publish this row's q in natural order for the y-neighbour lookup of the NEXT iteration
      __syncthreads();
#pragma unroll
      for (int r = 0; r < R1; ++r) cur[b + r * S1] = z[r];
      __syncthreads();
      if (ii > i0) {
        const float2* prv = rows_sm + ((ii - 1 - i0) & 1) * Gm::PITCH;
        const size_t ro = (size_t)(ii - 1) * ny;
        const float2* u2 = reinterpret_cast<const float2*>(u + ro);
        const float2* v2 = reinterpret_cast<const float2*>(v + ro);
        const float2* p2 = reinterpret_cast<const float2*>(p + ro);
        float2* uo = reinterpret_cast<float2*>(u_out + ro);
        float2* vo = reinterpret_cast<float2*>(v_out + ro);
        float2* po = reinterpret_cast<float2*>(p_out + ro);
        // all loads first: the outputs may alias the inputs (in-place projection), so the compiler
        // cannot hoist a load above an earlier store -- interleaving them costs R1 round trips per row
        float2 ua[R1], va[R1], pa[R1];
#pragma unroll
        for (int r = 0; r < R1; ++r) {
          const int n = b + r * S1;
          ua[r] = u2[n];
          va[r] = v2[n];
          pa[r] = p2[n];
        }
#pragma unroll
        for (int r = 0; r < R1; ++r) {
          const int n = b + r * S1;
          const float2 q0 = zprev[r], q1 = z[r];
          const float qn = prv[n + 1 < N ? n + 1 : 0].x;  // q[i-1][2n+2] (periodic)
          float2 uu = ua[r], vv = va[r];
          const float2 pp = pa[r];
          uu.x -= (q1.x - q0.x) * g.inv_dx;
          uu.y -= (q1.y - q0.y) * g.inv_dx;
          vv.x -= (q0.y - q0.x) * g.inv_dy;
          vv.y -= (qn - q0.y) * g.inv_dy;
          const float2 pn = make_float2(q0.x + pp.x, q0.y + pp.y);
          uo[n] = uu;
          vo[n] = vv;
          po[n] = pn;
          if (P.ho.halo) {  // slab mode: boundary rows are also the neighbours' halo rows (NVLink stores)
            const int im = ii - 1;
            if (im < P.ho.halo) {
              const size_t o = (size_t)im * ny;
              reinterpret_cast<float2*>(P.ho.lo[0] + o)[n] = uu;
              reinterpret_cast<float2*>(P.ho.lo[1] + o)[n] = vv;
              reinterpret_cast<float2*>(P.ho.lo[2] + o)[n] = pn;
            }
            if (im >= g.nx - P.ho.halo) {
              const size_t o = (size_t)(im - (g.nx - P.ho.halo)) * ny;
              reinterpret_cast<float2*>(P.ho.hi[0] + o)[n] = uu;
              reinterpret_cast<float2*>(P.ho.hi[1] + o)[n] = vv;
              reinterpret_cast<float2*>(P.ho.hi[2] + o)[n] = pn;
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R1; ++r) zprev[r] = z[r];
    }
    __syncthreads();
  }
}

// column pairs per CTA, shared memory per CTA and resident CTAs per SM of the column kernel
template <int R1, int R2, int R3>
void cols_geometry_t(int* pairs_out, size_t* smem_out, int* per_sm_out) {
  constexpr int N = R1 * R2 * R3;
  const size_t tables = (size_t)(N + R3 * (R2 + 1)) * sizeof(float2);
  const char* e = getenv("JAXIB_COL_PAIRS");
  int pairs = e ? atoi(e) : 4;  // column pairs (float4) per CTA
  if (pairs != 1 && pairs != 2 && pairs != 4) pairs = 4;
  while (pairs > 1 && (pairs * (N / R1) > 1024 || pairs * (size_t)(N + 2) * sizeof(float4) + tables > 220 * 1024)) pairs /= 2;
  const size_t smem = pairs * (size_t)(N + 2) * sizeof(float4) + tables;
  int per_sm = (int)((220 * 1024) / smem);
  const int by_threads = 2048 / (pairs * (N / R1));
  if (per_sm > by_threads) per_sm = by_threads;
  *pairs_out = pairs; *smem_out = smem; *per_sm_out = per_sm < 1 ? 1 : per_sm;
}

template <int R1, int R2, int R3>
int launch_cols_t(cudaStream_t s, const FastColParams& P, int batch, float2* spec) {
  constexpr int N = R1 * R2 * R3;
  int pairs, per_sm;
  size_t smem;
  cols_geometry_t<R1, R2, R3>(&pairs, &smem, &per_sm);
  dim3 grid((P.ncols + 2 * pairs - 1) / (2 * pairs), batch);
  FastColParams Q = P;
  {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    Q.first_wave = sms * per_sm;
  }
  // the opt-in for > 48 KB dynamic shared memory is set once per instantiation (idempotent; a
  // driver call on every launch costs several microseconds of host time on the critical path)
  static bool configured[5] = {false, false, false, false, false};
  if (!configured[pairs]) {
    const int cap = 220 * 1024;
    cudaFuncSetAttribute(cols_fast_kernel<R1, R2, R3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    cudaFuncSetAttribute(cols_fast_kernel<R1, R2, R3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    cudaFuncSetAttribute(cols_fast_kernel<R1, R2, R3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    configured[pairs] = true;
  }
  if (pairs == 4) {
    launch_pdl(16, cols_fast_kernel<R1, R2, R3, 4>, grid, dim3(4 * (N / R1)), smem, s, Q, spec);
  } else if (pairs == 2) {
    launch_pdl(16, cols_fast_kernel<R1, R2, R3, 2>, grid, dim3(2 * (N / R1)), smem, s, Q, spec);
  } else {
    launch_pdl(16, cols_fast_kernel<R1, R2, R3, 1>, grid, dim3(N / R1), smem, s, Q, spec);
  }
  return check_launch("cols_fast_kernel");
}

// ---- K5 for Nx = 8192: four passes, 1024 threads ----------------------------------------------------
// The three-pass instantiation (16, 32, 16) needs 128 registers for its radix-32 pass, i.e. 512 threads
// and 16 warps per SM (one 196 KB CTA): 0.43 MB/us against 0.74 MB/us of the 2048-row kernel.  With
// radices (8, 8, 8, 16) every butterfly fits the 64 registers of a 1024-thread CTA, at the price of one
// more trip through shared memory.  One column pair per CTA (128 KB); the pass-1 twiddles W_N^(gb r)
// are read from global memory (lane stride r * 8 bytes, r < 8: at most 14 sectors per warp load), the
// inner passes use small shared tables.  Position of frequency k after the forward passes:
//   (k % 8) * 1024 + ((k / 8) % 8) * 128 + ((k / 64) % 8) * 16 + k / 512   (= the plan's digit reversal).
template <int SW>
__device__ __forceinline__ int qswz_t(int p) { return p ^ ((p >> SW) & 7); }  // SW = log2(innermost radix)

template <int R1, int R2, int R3, int R4, int NP>
__global__ void __launch_bounds__(NP * R2 * R3 * R4, 1) cols4_fast_kernel(FastColParams P, float2* __restrict__ spec) {
  constexpr int N = R1 * R2 * R3 * R4;
  constexpr int S1 = N / R1;        // threads per column pair = stride of pass 1
  constexpr int L2 = N / R1;        // block length of pass 2
  constexpr int L3 = L2 / R2;       // block length of pass 3 = stride of pass 2
  constexpr int L4 = L3 / R3;       // = R4: block length of pass 4 = stride of pass 3
  constexpr int SW = R4 == 16 ? 4 : 3;
  constexpr int PQ = N + 2;         // pair pitch (see cols_fast_kernel)
  static_assert(L4 == R4 && (R4 == 8 || R4 == 16), "radix product");
  extern __shared__ float4 smq[];
  // inner passes: thread -> (pair, b) blocked; global passes: pair index fastest (see cols_fast_kernel)
  const int pair = threadIdx.x / S1, bt = threadIdx.x - pair * S1;
  const int gpair = threadIdx.x % NP, gb = threadIdx.x / NP;
  float2* const col2 = reinterpret_cast<float2*>(smq + pair * PQ);
  float4* const gcol = smq + gpair * PQ;
  float2* const tw2 = reinterpret_cast<float2*>(smq + NP * PQ);  // [L3][R2 + 1]: W_N^(b r R1)
  float2* const tw3 = tw2 + L3 * (R2 + 1);                       // [L4][R3 + 1]: W_N^(b r R1 R2)
  const int c = blockIdx.x * 2 * NP + 2 * pair;
  const int gc = blockIdx.x * 2 * NP + 2 * gpair;
  const bool gvalid = gc + 1 < P.limit;
  float4* g4 = reinterpret_cast<float4*>(spec + (size_t)blockIdx.y * P.nx * P.sp + gc);
  const size_t rs = (size_t)P.sp / 2;
  const float2* __restrict__ tw = P.tw;

  pdl_trigger();
  for (int e = threadIdx.x; e < L3 * R2; e += NP * S1) tw2[(e / R2) * (R2 + 1) + e % R2] = __ldg(tw + (e / R2) * (e % R2) * R1);
  for (int e = threadIdx.x; e < L4 * R3; e += NP * S1) tw3[(e / R3) * (R3 + 1) + e % R3] = __ldg(tw + (e / R3) * (e % R3) * R1 * R2);
  pdl_wait();
  {  // pass 1 forward: global -> registers -> shared
    float2 va[R1], vb[R1];
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gvalid) x = g4[(size_t)(gb + r * S1) * rs];
      va[r] = make_float2(x.x, x.y);
      vb[r] = make_float2(x.z, x.w);
    }
    Dft<R1>::run(va);
    Dft<R1>::run(vb);
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      if (r > 0) {
        const float2 w = __ldg(tw + gb * r);
        va[r] = cmul(va[r], w);
        vb[r] = cmul(vb[r], w);
      }
      gcol[qswz_t<SW>(gb + r * S1)] = make_float4(va[r].x, va[r].y, vb[r].x, vb[r].y);
    }
  }
  __syncthreads();
  // pass 2 forward (radix R2, stride L3), one column per thread
#pragma unroll
  for (int m = 0; m < 2 * (N / R2) / S1; ++m) {
    const int w = bt + m * S1;
    const int cc = w & 1, bf = w >> 1;
    const int blk = bf / L3, b = bf - blk * L3;
    const int p0 = blk * L2 + b;
    float2 v[R2];
#pragma unroll
    for (int r = 0; r < R2; ++r) v[r] = col2[2 * qswz_t<SW>(p0 + r * L3) + cc];
    Dft<R2>::run(v);
#pragma unroll
    for (int r = 1; r < R2; ++r) v[r] = cmul(v[r], tw2[b * (R2 + 1) + r]);
#pragma unroll
    for (int r = 0; r < R2; ++r) col2[2 * qswz_t<SW>(p0 + r * L3) + cc] = v[r];
  }
  __syncthreads();
  // pass 3 forward (radix R3, stride L4)
#pragma unroll
  for (int m = 0; m < 2 * (N / R3) / S1; ++m) {
    const int w = bt + m * S1;
    const int cc = w & 1, bf = w >> 1;
    const int blk = bf / L4, b = bf - blk * L4;
    const int p0 = blk * L3 + b;
    float2 v[R3];
#pragma unroll
    for (int r = 0; r < R3; ++r) v[r] = col2[2 * qswz_t<SW>(p0 + r * L4) + cc];
    Dft<R3>::run(v);
#pragma unroll
    for (int r = 1; r < R3; ++r) v[r] = cmul(v[r], tw3[b * (R3 + 1) + r]);
#pragma unroll
    for (int r = 0; r < R3; ++r) col2[2 * qswz_t<SW>(p0 + r * L4) + cc] = v[r];
  }
  __syncthreads();
  // pass 4 forward, diagonal, pass 4 inverse -- in registers
#pragma unroll
  for (int m = 0; m < (2 * (N / R4) + S1 - 1) / S1; ++m) {
    const int w = bt + m * S1;
    if (w >= 2 * (N / R4)) continue;
    const int cc = w & 1, p0 = (w >> 1) * R4;
    const float ly = P.lam_y[c + cc];
    float2 v[R4];
#pragma unroll
    for (int r = 0; r < R4; ++r) v[r] = col2[2 * qswz_t<SW>(p0 + r) + cc];
    Dft<R4>::run(v);
    const float4* lx4 = reinterpret_cast<const float4*>(P.lam_x_pos + p0);
#pragma unroll
    for (int r = 0; r < R4; r += 4) {
      const float4 l4 = __ldg(lx4 + r / 4);
      v[r] = cscale(v[r], inv_eig_f(l4.x + ly));
      v[r + 1] = cscale(v[r + 1], inv_eig_f(l4.y + ly));
      v[r + 2] = cscale(v[r + 2], inv_eig_f(l4.z + ly));
      v[r + 3] = cscale(v[r + 3], inv_eig_f(l4.w + ly));
    }
    dft_dir<R4, true>(v);
#pragma unroll
    for (int r = 0; r < R4; ++r) col2[2 * qswz_t<SW>(p0 + r) + cc] = v[r];
  }
  __syncthreads();
  // pass 3 inverse
#pragma unroll
  for (int m = 0; m < 2 * (N / R3) / S1; ++m) {
    const int w = bt + m * S1;
    const int cc = w & 1, bf = w >> 1;
    const int blk = bf / L4, b = bf - blk * L4;
    const int p0 = blk * L3 + b;
    float2 v[R3];
#pragma unroll
    for (int r = 0; r < R3; ++r) v[r] = col2[2 * qswz_t<SW>(p0 + r * L4) + cc];
#pragma unroll
    for (int r = 1; r < R3; ++r) v[r] = cmulc(v[r], tw3[b * (R3 + 1) + r]);
    dft_dir<R3, true>(v);
#pragma unroll
    for (int r = 0; r < R3; ++r) col2[2 * qswz_t<SW>(p0 + r * L4) + cc] = v[r];
  }
  __syncthreads();
  // pass 2 inverse
#pragma unroll
  for (int m = 0; m < 2 * (N / R2) / S1; ++m) {
    const int w = bt + m * S1;
    const int cc = w & 1, bf = w >> 1;
    const int blk = bf / L3, b = bf - blk * L3;
    const int p0 = blk * L2 + b;
    float2 v[R2];
#pragma unroll
    for (int r = 0; r < R2; ++r) v[r] = col2[2 * qswz_t<SW>(p0 + r * L3) + cc];
#pragma unroll
    for (int r = 1; r < R2; ++r) v[r] = cmulc(v[r], tw2[b * (R2 + 1) + r]);
    dft_dir<R2, true>(v);
#pragma unroll
    for (int r = 0; r < R2; ++r) col2[2 * qswz_t<SW>(p0 + r * L3) + cc] = v[r];
  }
  __syncthreads();
  {  // pass 1 inverse: shared -> registers -> global
    float2 va[R1], vb[R1];
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      const float4 x = gcol[qswz_t<SW>(gb + r * S1)];
      va[r] = make_float2(x.x, x.y);
      vb[r] = make_float2(x.z, x.w);
      if (r > 0) {
        const float2 w = __ldg(tw + gb * r);
        va[r] = cmulc(va[r], w);
        vb[r] = cmulc(vb[r], w);
      }
    }
    dft_dir<R1, true>(va);
    dft_dir<R1, true>(vb);
    if (gvalid) {
#pragma unroll
      for (int r = 0; r < R1; ++r)
        g4[(size_t)(gb + r * S1) * rs] = make_float4(va[r].x, va[r].y, vb[r].x, vb[r].y);
    }
  }
}

template <int R1, int R2, int R3, int R4, int NP>
int launch_cols4_t(cudaStream_t s, const FastColParams& P, int batch, float2* spec) {
  constexpr int N = R1 * R2 * R3 * R4;
  const size_t smem = (size_t)NP * (N + 2) * sizeof(float4) +
                      ((size_t)(N / R1 / R2) * (R2 + 1) + (size_t)R4 * (R3 + 1)) * sizeof(float2);
  dim3 grid((P.ncols + 2 * NP - 1) / (2 * NP), batch);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(cols4_fast_kernel<R1, R2, R3, R4, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    configured = true;
  }
  launch_pdl(16, cols4_fast_kernel<R1, R2, R3, R4, NP>, grid, dim3(NP * (N / R1)), smem, s, P, spec);
  return check_launch("cols4_fast_kernel");
}

static bool four_pass_8192() {
  static int on = -1;
  if (on < 0) on = getenv("JAXIB_COLS8192_3PASS") == nullptr ? 1 : 0;  // =1: the 512-thread three-pass kernel (A/B)
  return on == 1;
}

template <int R1, int R2, int R3>
int launch_rows_fwd_t(cudaStream_t s, const FastRowParams& P, const float* u, const float* v, float2* spec) {
  dim3 grid((P.g.nx + P.rows_per_group - 1) / P.rows_per_group, P.g.batch);
  const size_t smem = (RowGeom<R1, R2, R3>::PITCH + 1 + R3 * (R2 + 1)) * sizeof(float2) + 3 * (size_t)P.g.ny * sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(rows_fwd_fast_kernel<R1, R2, R3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured = true;
  }
  launch_pdl(8, rows_fwd_fast_kernel<R1, R2, R3>, grid, dim3(R2 * R3), smem, s, P, u, v, spec);
  return check_launch("rows_fwd_fast_kernel");
}

template <int R1, int R2, int R3>
int launch_rows_inv_t(cudaStream_t s, const FastRowParams& P, const float2* spec, const float* u, const float* v,
                      const float* p, float* uo, float* vo, float* po, bool project) {
  dim3 grid((P.g.nx + P.rows_per_group - 1) / P.rows_per_group, P.g.batch);
  const size_t smem = (2 * RowGeom<R1, R2, R3>::PITCH + R1 * R2 * R3 + 4 + R3 * (R2 + 1)) * sizeof(float2);
  static bool configured = false;
  if (!configured) {  // opt in to > 48 KB dynamic shared memory once (idempotent)
    cudaFuncSetAttribute(rows_inv_fast_kernel<R1, R2, R3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(rows_inv_fast_kernel<R1, R2, R3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured = true;
  }
  // measured: behind the 8192-row column kernel (one 196 KB CTA per SM) an early-launched K6 costs 5 % of
  // the step; everywhere else (smaller grids, slab mode: K6 follows the barrier kernel) it gains 1-6 %
  const int pdl_bit = (P.g.nx_global == 0 && P.g.nx >= 8192) ? 0 : 32;
  if (project)
    launch_pdl(pdl_bit, rows_inv_fast_kernel<R1, R2, R3, true>, grid, dim3(R2 * R3), smem, s, P, spec, u, v, p, uo, vo, po);
  else
    launch_pdl(pdl_bit, rows_inv_fast_kernel<R1, R2, R3, false>, grid, dim3(R2 * R3), smem, s, P, spec, u, v, p, uo, vo, po);
  return check_launch("rows_inv_fast_kernel");
}

}  // namespace

bool fast_size_supported(int n) { return n == 512 || n == 1024 || n == 2048 || n == 4096; }
bool fast_cols_supported(int n) { return fast_size_supported(n) || n == 8192; }

void fast_radices(int n, int* r) {
  switch (n) {
    case 512: r[0] = 8; r[1] = 8; r[2] = 8; break;
    case 1024: r[0] = 8; r[1] = 8; r[2] = 16; break;
    case 2048: r[0] = 8; r[1] = 16; r[2] = 16; break;
    case 4096: r[0] = 16; r[1] = 16; r[2] = 16; break;  // (8, 8, 8, 8) with 2 pairs per CTA measured within 1 %
    case 8192:  // columns only
      if (four_pass_8192()) { r[0] = 8; r[1] = 8; r[2] = 8; r[3] = 16; }
      else { r[0] = 16; r[1] = 32; r[2] = 16; r[3] = 0; }
      break;
    default: r[0] = r[1] = r[2] = 0;
  }
  if (n != 8192) r[3] = 0;
}

// columns: must agree with fast_radices() (the eigenvalue table is stored in this digit order)
#define JAXIB_DISPATCH_COLS(n, CALL)          \
  switch (n) {                                \
    case 512: return CALL(8, 8, 8);           \
    case 1024: return CALL(8, 8, 16);         \
    case 2048: return CALL(8, 16, 16);        \
    case 4096: return CALL(16, 16, 16);       \
    case 8192: return CALL(16, 32, 16);       \
    default: set_error("fast FFT path: unsupported size %d", n); return JAXIB_ERR_UNSUPPORTED; \
  }

void fast_cols_geometry(int nx, int* cols_per_cta, int* ctas_per_sm) {
  int pairs = 1, per_sm = 1;
  size_t smem = 0;
  switch (nx) {
    case 512: cols_geometry_t<8, 8, 8>(&pairs, &smem, &per_sm); break;
    case 1024: cols_geometry_t<8, 8, 16>(&pairs, &smem, &per_sm); break;
    case 2048: cols_geometry_t<8, 16, 16>(&pairs, &smem, &per_sm); break;
    case 4096: cols_geometry_t<16, 16, 16>(&pairs, &smem, &per_sm); break;
    case 8192: cols_geometry_t<16, 32, 16>(&pairs, &smem, &per_sm); break;  // one pair per CTA, one CTA per SM: both kernels
    default: break;
  }
  *cols_per_cta = 2 * pairs;
  *ctas_per_sm = per_sm;
}

int launch_cols_fast(cudaStream_t s, const FastColParams& P, int batch, float2* spec) {
  if (P.nx == 8192 && four_pass_8192()) return launch_cols4_t<8, 8, 8, 16, 1>(s, P, batch, spec);
#define CALL_COLS(a, b, c) launch_cols_t<a, b, c>(s, P, batch, spec)
  JAXIB_DISPATCH_COLS(P.nx, CALL_COLS)
#undef CALL_COLS
}

// rows: first radix 8 where possible = twice the threads per row (the row kernels are latency bound)
#define JAXIB_DISPATCH_ROWS(n, CALL)          \
  switch (n) {                                \
    case 512: return CALL(8, 8, 8);           \
    case 1024: return CALL(8, 16, 8);         \
    case 2048: return CALL(8, 16, 16);        \
    case 4096: return CALL(16, 16, 16);       \
    default: set_error("fast FFT path: unsupported size %d", n); return JAXIB_ERR_UNSUPPORTED; \
  }

// resident CTAs per SM of the row kernels for this row length (wave-quantisation model in plan_create)

template <int R1, int R2, int R3>
int rows_ctas_per_sm_t(int ny, bool inverse) {
  int nb = 1;
  if (inverse) {
    const size_t smem = (2 * RowGeom<R1, R2, R3>::PITCH + R1 * R2 * R3 + 4 + R3 * (R2 + 1)) * sizeof(float2);
    cudaFuncSetAttribute(rows_inv_fast_kernel<R1, R2, R3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rows_inv_fast_kernel<R1, R2, R3, true>, R2 * R3, smem);
  } else {
    const size_t smem = (RowGeom<R1, R2, R3>::PITCH + 1 + R3 * (R2 + 1)) * sizeof(float2) + 3 * (size_t)ny * sizeof(float);
    cudaFuncSetAttribute(rows_fwd_fast_kernel<R1, R2, R3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rows_fwd_fast_kernel<R1, R2, R3>, R2 * R3, smem);
  }
  return nb < 1 ? 1 : nb;
}

int fast_rows_ctas_per_sm(int ny, bool inverse) {
  switch (ny / 2) {
    case 512: return rows_ctas_per_sm_t<8, 8, 8>(ny, inverse);
    case 1024: return rows_ctas_per_sm_t<8, 16, 8>(ny, inverse);
    case 2048: return rows_ctas_per_sm_t<8, 16, 16>(ny, inverse);
    case 4096: return rows_ctas_per_sm_t<16, 16, 16>(ny, inverse);
    default: return 1;
  }
}

int launch_rows_fwd_fast(cudaStream_t s, const FastRowParams& P, const float* u, const float* v, float2* spec) {
#define CALL_RF(a, b, c) launch_rows_fwd_t<a, b, c>(s, P, u, v, spec)
  JAXIB_DISPATCH_ROWS(P.g.ny / 2, CALL_RF)
#undef CALL_RF
}

int launch_rows_inv_fast(cudaStream_t s, const FastRowParams& P, const float2* spec, const float* u, const float* v,
                         const float* p, float* uo, float* vo, float* po, bool project) {
#define CALL_RI(a, b, c) launch_rows_inv_t<a, b, c>(s, P, spec, u, v, p, uo, vo, po, project)
  JAXIB_DISPATCH_ROWS(P.g.ny / 2, CALL_RI)
#undef CALL_RI
}

}  // namespace jaxib
